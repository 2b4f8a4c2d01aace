#!/bin/bash
# Round 2, call N (1 GPU, ~3 min): attention with the barrier polls issued early (MMA warp and softmax warps).
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
run() { name=$1; shift; local t0=$SECONDS; timeout "$@" > gpurun_out/$name.log 2>&1; local rc=$?; echo "$name exit $rc ($((SECONDS - t0)) s)" >> gpurun_out/summary.txt; return $rc; }
PT="python -m pytest -q -m gpu --no-header -p no:cacheprovider --tb=short -x"
run attn_tests 200 $PT tests/test_kernels_gpu.py -k "attention or mask"
run attn_bench 100 python tools/attn_bench.py
run attn_bench_cfg3 100 python tools/attn_bench.py 32 4 256 256
run attn_bench_cfg5 100 python tools/attn_bench.py 4 4 512 512
VGPT_ATTN_VARIANT=8 run attn_trace 100 python tools/attn_trace.py
run sp_tests 300 $PT tests/test_sequence_parallel.py tests/test_zz_batch_gpu.py tests/test_model_gpu.py
run bench_cfg2 200 python bench.py --steps 3 --warmup 3 --no-baselines --strong none
for f in attn_tests attn_bench attn_bench_cfg3 attn_bench_cfg5 sp_tests bench_cfg2; do echo "=== $f"; tail -n ${TAILN:-6} gpurun_out/$f.log 2>/dev/null | cut -c1-400; done
cat gpurun_out/summary.txt
