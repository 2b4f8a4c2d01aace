#!/bin/bash
# Round-2 first contact (1 GPU, ~20 min of box time): everything that was written after the round-1
# GPU minutes were spent gets validated and measured in ONE call.  Every stage under its own timeout;
# logs and reports into gpurun_out/.   gpurun --timeout 1500 -- 'bash tools/gpu/r02_first.sh'
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
run() { name=$1; shift; timeout "$@" > gpurun_out/$name.log 2>&1; echo "$name exit $?" >> gpurun_out/summary.txt; }
PT="python -m pytest -q -m gpu --no-header -p no:cacheprovider --tb=short"
# 1. the whole parity suite (includes tests/test_zz_*_gpu.py: batch, rollout -- first hardware run)
run tests        1500 $PT tests
# 2. experimental kernels, opt-in: skinny-tail GEMM bit-exactness, then its effect on the bench line
VGPT_TEST_EXPERIMENTAL=1 run skinny_test 300 $PT tests/test_kernels_gpu.py -k skinny
VGPT_TEST_EXPERIMENTAL=1 run attn_variants 300 $PT tests/test_kernels_gpu.py -k variants
for v in 0 1 2 3 4 5; do VGPT_ATTN_VARIANT=$v run attnbench_var$v 200 python tools/attn_bench.py; done
VGPT_ATTN_VARIANT=8 run attn_trace 200 python tools/attn_trace.py
VGPT_ATTN_VARIANT=3 run tests_var3 900 $PT tests/test_kernels_gpu.py tests/test_model_gpu.py -k "attention or next_clip"
run smoke        200 python __graft_entry__.py --smoke
run bench_cfg2   600 python bench.py --steps 3 --warmup 3
VGPT_GEMM_SKINNY_TAIL=1 run bench_cfg2_skinny 600 python bench.py --steps 3 --warmup 3
VGPT_ATTN_VARIANT=3 run bench_cfg2_var3 600 python bench.py --steps 3 --warmup 3
run gemmsweep    300 python tools/gemm_bench.py
VGPT_GEMM_SKINNY_TAIL=1 run gemmsweep_skinny 300 python tools/gemm_bench.py
# 3. workloads that have never been measured: batch of videos (4 per pass), rollout with / without the cache
run bench_cfg4   900 python bench.py --config cfg4 --steps 1 --warmup 3 --batch 4
run bench_cfg4_b8 900 python bench.py --config cfg4 --steps 1 --warmup 3 --batch 8
run bench_roll   600 python bench.py --config cfg3 --rollout 3 --steps 1 --warmup 3
run bench_roll_re 600 python bench.py --config cfg3 --rollout 3 --recompute --steps 1 --warmup 3
# 4. profiles of the CURRENT kernels: launch list, then --set full of the four GEMM shapes and attention
python tools/profile_step.py > gpurun_out/plain.log 2>&1 &&
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/launches.csv python tools/profile_step.py > gpurun_out/ncu_launches.log 2>&1
echo "launches exit $?" >> gpurun_out/summary.txt
python tools/profile_step.py --no-prefill > gpurun_out/plain2.log 2>&1 &&
ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:gemm_bf16 -s 4 -c 4 \
    -o gpurun_out/prof_gemm -f python tools/profile_step.py --no-prefill > gpurun_out/ncu_gemm.log 2>&1
echo "ncu_gemm exit $?" >> gpurun_out/summary.txt
ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:attn_pair -s 1 -c 1 \
    -o gpurun_out/prof_attn -f python tools/profile_step.py --no-prefill > gpurun_out/ncu_attn.log 2>&1
echo "ncu_attn exit $?" >> gpurun_out/summary.txt
for f in tests skinny_test attn_variants attnbench_var0 attnbench_var1 attnbench_var2 attnbench_var3 attnbench_var4 attnbench_var5 tests_var3 bench_cfg2_var3 smoke bench_cfg2 bench_cfg2_skinny bench_cfg4 bench_cfg4_b8 bench_roll bench_roll_re; do
  echo "=== $f"; tail -n ${TAILN:-6} gpurun_out/$f.log | cut -c1-600; done
cat gpurun_out/summary.txt
