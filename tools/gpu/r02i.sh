#!/bin/bash
# Round 2, call I (1 GPU, ~4 min): attention with one MMA issuer per query tile.
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
run() { name=$1; shift; local t0=$SECONDS; timeout "$@" > gpurun_out/$name.log 2>&1; local rc=$?; echo "$name exit $rc ($((SECONDS - t0)) s)" >> gpurun_out/summary.txt; return $rc; }
PT="python -m pytest -q -m gpu --no-header -p no:cacheprovider --tb=short -x"
run attn_tests 200 $PT tests/test_kernels_gpu.py -k "attention or mask"
run attn_bench 100 python tools/attn_bench.py
run attn_bench_cfg3 100 python tools/attn_bench.py 32 4 256 256
run attn_bench_cfg5 100 python tools/attn_bench.py 4 4 512 512
run attn_bench_d128 100 python tools/attn_bench.py 4 4 256 256 24 128
VGPT_ATTN_VARIANT=8 run attn_trace 100 python tools/attn_trace.py
run model_tests 600 $PT tests/test_model_gpu.py tests/test_zz_batch_gpu.py tests/test_zz_rollout_gpu.py tests/test_sequence_parallel.py
run bench_cfg2 200 python bench.py --steps 3 --warmup 3 --no-baselines --strong none
timeout 300 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:attn_pair -s 40 -c 1 \
    -o gpurun_out/prof_attn -f python tools/profile_step.py > gpurun_out/ncu_attn.log 2>&1
echo "ncu_attn exit $?" >> gpurun_out/summary.txt
for f in attn_tests attn_bench attn_bench_cfg3 attn_bench_cfg5 attn_bench_d128 model_tests bench_cfg2; do
  echo "=== $f"; tail -n ${TAILN:-8} gpurun_out/$f.log 2>/dev/null | cut -c1-330; done
cat gpurun_out/summary.txt
