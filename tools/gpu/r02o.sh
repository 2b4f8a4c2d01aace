#!/bin/bash
# Round 2, call O (1 GPU, ~3 min): attention with two softmax warps per row block (VGPT_ATTN_HALVES=2) against the default.
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
run() { name=$1; shift; local t0=$SECONDS; timeout "$@" > gpurun_out/$name.log 2>&1; local rc=$?; echo "$name exit $rc ($((SECONDS - t0)) s)" >> gpurun_out/summary.txt; return $rc; }
PT="python -m pytest -q -m gpu --no-header -p no:cacheprovider --tb=short -x"
export VGPT_ATTN_HALVES=2
if run attn_tests_h2 150 $PT tests/test_kernels_gpu.py -k "attention or mask"; then
  run attn_bench_h2 60 python tools/attn_bench.py
  run attn_bench_cfg3_h2 60 python tools/attn_bench.py 32 4 256 256
  run attn_bench_cfg5_h2 60 python tools/attn_bench.py 4 4 512 512
  run attn_bench_d128_h2 60 python tools/attn_bench.py 4 4 256 256 24 128
  run sp_tests_h2 300 $PT tests/test_sequence_parallel.py tests/test_zz_batch_gpu.py tests/test_model_gpu.py
  run bench_cfg2_h2 200 python bench.py --steps 3 --warmup 3 --no-baselines --strong none
fi
unset VGPT_ATTN_HALVES
run attn_bench_h1 60 python tools/attn_bench.py
run attn_bench_cfg5_h1 60 python tools/attn_bench.py 4 4 512 512
run bench_cfg2_h1 200 python bench.py --steps 3 --warmup 3 --no-baselines --strong none
for f in attn_tests_h2 attn_bench_h2 attn_bench_cfg3_h2 attn_bench_cfg5_h2 attn_bench_d128_h2 sp_tests_h2 bench_cfg2_h2 attn_bench_h1 attn_bench_cfg5_h1 bench_cfg2_h1; do echo "=== $f"; tail -n ${TAILN:-6} gpurun_out/$f.log 2>/dev/null | cut -c1-300; done
cat gpurun_out/summary.txt
