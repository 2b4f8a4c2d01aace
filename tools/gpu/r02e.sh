#!/bin/bash
# Round 2, call E (1 GPU, ~10 min): tail MMA at N = 32 + one W box per special piece, fused final kernel, new small
# kernels, new parity tests; launch list of a step.
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
run() { name=$1; shift; local t0=$SECONDS; timeout "$@" > gpurun_out/$name.log 2>&1; echo "$name exit $? ($((SECONDS - t0)) s)" >> gpurun_out/summary.txt; }
PT="python -m pytest -q -m gpu --no-header -p no:cacheprovider --tb=short"
run kernel_tests 600 $PT tests/test_kernels_gpu.py tests/test_umma_layouts.py
run umma_rate 100 python tools/umma_rate.py
VGPT_ATTN_VARIANT=8 run attn_trace 100 python tools/attn_trace.py
run gemmsweep 120 python tools/gemm_bench.py
run gemmsweep_1040 120 python tools/gemm_bench.py 1040
run model_tests 900 $PT tests/test_model_gpu.py tests/test_zz_batch_gpu.py tests/test_zz_rollout_gpu.py tests/test_sequence_parallel.py
run bench_cfg2 300 python bench.py --steps 3 --warmup 3 --no-baselines --strong none
VGPT_GEMM_TAIL_IN_LOOP=0 run bench_cfg2_plain 300 python bench.py --steps 3 --warmup 3 --no-baselines --strong none
run smoke 200 python __graft_entry__.py --smoke
python tools/profile_step.py > gpurun_out/plain.log 2>&1 &&
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/launches.csv python tools/profile_step.py > gpurun_out/ncu_launches.log 2>&1
echo "launches exit $?" >> gpurun_out/summary.txt
for f in kernel_tests model_tests gemmsweep gemmsweep_1040 bench_cfg2 bench_cfg2_plain smoke; do
  echo "=== $f"; tail -n ${TAILN:-22} gpurun_out/$f.log | cut -c1-400; done
echo "=== umma_rate"; tail -3 gpurun_out/umma_rate.log
cat gpurun_out/summary.txt
