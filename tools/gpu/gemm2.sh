#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
run() { name=$1; shift; timeout "$@" > gpurun_out/$name.log 2>&1; echo "$name exit $?" >> gpurun_out/summary.txt; }
PT="python -m pytest -q -m gpu --no-header -p no:cacheprovider --tb=short"
run gemm      400 $PT tests/test_kernels_gpu.py -k "gemm"
run sweep_real 300 python tools/gemm_bench.py
for f in gemm sweep_real; do echo "=== $f"; tail -n 20 gpurun_out/$f.log; done
cat gpurun_out/summary.txt
