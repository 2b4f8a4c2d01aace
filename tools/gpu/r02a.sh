#!/bin/bash
# Round 2, call A (1 GPU, ~30 min): cfg3 / cfg5 on one GPU first (never measured, cfg3 hung under SP in round 1),
# then everything written after the round-1 GPU minutes were spent, the 50-step parity floor, cfg4, and ncu
# captures of the production kernels.  Every stage under its own timeout; logs into gpurun_out/.
#   gpurun --timeout 2700 -- 'bash tools/gpu/r02a.sh'
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
run() { name=$1; shift; local t0=$SECONDS; timeout "$@" > gpurun_out/$name.log 2>&1; echo "$name exit $? ($((SECONDS - t0)) s)" >> gpurun_out/summary.txt; }
PT="python -m pytest -q -m gpu --no-header -p no:cacheprovider --tb=short"
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > gpurun_out/smi.txt 2>&1
# 1. long context on one GPU
run attn_cfg3_test 300 $PT tests/test_kernels_gpu.py -k cfg3_geometry
run attnbench_cfg3 200 python tools/attn_bench.py 32 4 256 256
VGPT_FAULT_DUMP=250 run bench_cfg3 300 python bench.py --config cfg3 --steps 2 --warmup 3 --no-baselines
VGPT_FAULT_DUMP=250 run bench_cfg5 300 python bench.py --config cfg5 --steps 2 --warmup 3 --no-baselines
# 2. opt-in kernels: bit-exactness / tolerance tests, then timings
VGPT_TEST_EXPERIMENTAL=1 run skinny_test 300 $PT tests/test_kernels_gpu.py -k skinny
VGPT_TEST_EXPERIMENTAL=1 run attn_variants 300 $PT tests/test_kernels_gpu.py -k variants
VGPT_TEST_EXPERIMENTAL=1 run fold_test 300 $PT tests/test_model_gpu.py -k folding
for v in 0 1 2 3 4 5; do VGPT_ATTN_VARIANT=$v run attnbench_var$v 200 python tools/attn_bench.py; done
VGPT_ATTN_VARIANT=8 run attn_trace 200 python tools/attn_trace.py
run gemmsweep    300 python tools/gemm_bench.py
VGPT_GEMM_SKINNY_TAIL=1 run gemmsweep_skinny 300 python tools/gemm_bench.py
run umma_rate    200 python tools/umma_rate.py
# 3. the headline line with the default kernels, with the skinny tail, with folded norms, with attention variants
run bench_cfg2   600 python bench.py --steps 3 --warmup 3
VGPT_GEMM_SKINNY_TAIL=1 run bench_cfg2_skinny 300 python bench.py --steps 2 --warmup 3 --no-baselines
VGPT_FOLD_RMSNORM=1 run bench_cfg2_fold 300 python bench.py --steps 2 --warmup 3 --no-baselines
VGPT_ATTN_VARIANT=3 VGPT_GEMM_SKINNY_TAIL=1 run bench_cfg2_var3_skinny 300 python bench.py --steps 2 --warmup 3 --no-baselines
# 4. parity: 50-step full size, bf16-vs-bf16 floor, per-layer errors
run parity_floor 900 python tools/parity_floor.py --out gpurun_out/parity_floor.json
# 5. configs[3]
run bench_cfg4   600 python bench.py --config cfg4 --steps 1 --warmup 3 --batch 4 --no-baselines
# 6. ncu: launch list, --set full of the production GEMMs and attention
python tools/profile_step.py > gpurun_out/plain.log 2>&1 &&
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/launches.csv python tools/profile_step.py > gpurun_out/ncu_launches.log 2>&1
echo "launches exit $?" >> gpurun_out/summary.txt
timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:gemm_bf16 -s 4 -c 4 \
    -o gpurun_out/prof_gemm -f python tools/profile_step.py --no-prefill > gpurun_out/ncu_gemm.log 2>&1
echo "ncu_gemm exit $?" >> gpurun_out/summary.txt
timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:attn_pair -s 1 -c 1 \
    -o gpurun_out/prof_attn -f python tools/profile_step.py --no-prefill > gpurun_out/ncu_attn.log 2>&1
echo "ncu_attn exit $?" >> gpurun_out/summary.txt
for f in attn_cfg3_test attnbench_cfg3 bench_cfg3 bench_cfg5 skinny_test attn_variants fold_test attnbench_var0 attnbench_var1 attnbench_var2 \
         attnbench_var3 attnbench_var4 attnbench_var5 bench_cfg2 bench_cfg2_skinny bench_cfg2_fold bench_cfg2_var3_skinny parity_floor bench_cfg4; do
  echo "=== $f"; tail -n ${TAILN:-4} gpurun_out/$f.log | cut -c1-400; done
cat gpurun_out/summary.txt
