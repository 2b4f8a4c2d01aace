#!/bin/bash
# Round 2, call R (1 GPU, ~3 min): residual epilogue with the residual requested one chunk ahead.
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
run() { name=$1; shift; local t0=$SECONDS; timeout "$@" > gpurun_out/$name.log 2>&1; local rc=$?; echo "$name exit $rc ($((SECONDS - t0)) s)" >> gpurun_out/summary.txt; return $rc; }
PT="python -m pytest -q -m gpu --no-header -p no:cacheprovider --tb=short -x"
run gemm_tests 300 $PT tests/test_kernels_gpu.py -k "gemm"
run gemmsweep 120 python tools/gemm_bench.py
run model_tests 300 $PT tests/test_model_gpu.py tests/test_sequence_parallel.py tests/test_zz_batch_gpu.py
run bench_cfg2 200 python bench.py --steps 3 --warmup 3 --no-baselines --strong none
for f in gemm_tests gemmsweep model_tests bench_cfg2; do echo "=== $f"; tail -n ${TAILN:-22} gpurun_out/$f.log 2>/dev/null | cut -c1-300; done
cat gpurun_out/summary.txt
