#!/bin/bash
# Round-2 first contact, part 2 (1 GPU, ~15 min): workloads that have never been measured, and --set full
# captures of the CURRENT kernels (the roofline.traffic figure in bench.py still quotes an r01c capture).
#   gpurun --timeout 1200 -- 'bash tools/gpu/r02_second.sh'
mkdir -p gpurun_out; rm -f gpurun_out/summary2.txt
run() { name=$1; shift; timeout "$@" > gpurun_out/$name.log 2>&1; echo "$name exit $?" >> gpurun_out/summary2.txt; }
run bench_cfg4    900 python bench.py --config cfg4 --steps 1 --warmup 3 --batch 4
run bench_cfg4_b8 900 python bench.py --config cfg4 --steps 1 --warmup 3 --batch 8
run bench_roll    600 python bench.py --config cfg3 --rollout 3 --steps 1 --warmup 3
run bench_roll_re 600 python bench.py --config cfg3 --rollout 3 --recompute --steps 1 --warmup 3
python tools/profile_step.py > gpurun_out/plain.log 2>&1 &&
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/launches.csv python tools/profile_step.py > gpurun_out/ncu_launches.log 2>&1
echo "launches exit $?" >> gpurun_out/summary2.txt
python tools/profile_step.py --no-prefill > gpurun_out/plain2.log 2>&1 &&
ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:gemm_bf16 -s 4 -c 4 \
    -o gpurun_out/prof_gemm -f python tools/profile_step.py --no-prefill > gpurun_out/ncu_gemm.log 2>&1
echo "ncu_gemm exit $?" >> gpurun_out/summary2.txt
ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:attn_pair -s 1 -c 1 \
    -o gpurun_out/prof_attn -f python tools/profile_step.py --no-prefill > gpurun_out/ncu_attn.log 2>&1
echo "ncu_attn exit $?" >> gpurun_out/summary2.txt
for f in bench_cfg4 bench_cfg4_b8 bench_roll bench_roll_re; do echo "=== $f"; tail -n 3 gpurun_out/$f.log | cut -c1-700; done
cat gpurun_out/summary2.txt
