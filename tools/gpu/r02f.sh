#!/bin/bash
# Round 2, call F (1 GPU, ~10 min): attention predicate in registers + lean MMA-warp descriptors, split-K o / down;
# first 1-GPU lines of cfg3 / cfg5 / cfg4.
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
run() { name=$1; shift; local t0=$SECONDS; timeout "$@" > gpurun_out/$name.log 2>&1; echo "$name exit $? ($((SECONDS - t0)) s)" >> gpurun_out/summary.txt; }
PT="python -m pytest -q -m gpu --no-header -p no:cacheprovider --tb=short -x"
run kernel_tests 600 $PT tests/test_kernels_gpu.py tests/test_umma_layouts.py
run attn_bench 100 python tools/attn_bench.py
VGPT_ATTN_VARIANT=8 run attn_trace 100 python tools/attn_trace.py
run gemmsweep 150 python tools/gemm_bench.py
run gemmsweep_1040 150 python tools/gemm_bench.py 1040
run model_tests 900 $PT tests/test_model_gpu.py tests/test_zz_batch_gpu.py tests/test_zz_rollout_gpu.py tests/test_sequence_parallel.py
run bench_cfg2 300 python bench.py --steps 3 --warmup 3 --no-baselines --strong none
VGPT_GEMM_SPLITK=0 run bench_cfg2_nosplit 300 python bench.py --steps 3 --warmup 3 --no-baselines --strong none
run bench_cfg3 300 python bench.py --config cfg3 --steps 2 --warmup 3 --no-baselines --strong none
run bench_cfg5 300 python bench.py --config cfg5 --steps 2 --warmup 3 --no-baselines --strong none
run bench_cfg4_b4 400 python bench.py --config cfg4 --batch 4 --videos 8 --steps 1 --warmup 3 --no-baselines --strong none
run smoke 200 python __graft_entry__.py --smoke
python tools/profile_step.py > gpurun_out/plain.log 2>&1 &&
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/launches.csv python tools/profile_step.py > gpurun_out/ncu_launches.log 2>&1
echo "launches exit $?" >> gpurun_out/summary.txt
for f in kernel_tests attn_bench model_tests gemmsweep gemmsweep_1040 bench_cfg2 bench_cfg2_nosplit bench_cfg3 bench_cfg5 bench_cfg4_b4 smoke; do
  echo "=== $f"; tail -n ${TAILN:-24} gpurun_out/$f.log | cut -c1-330; done
cat gpurun_out/summary.txt
