#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
run() { name=$1; shift; timeout "$@" > gpurun_out/$name.log 2>&1; echo "$name exit $?" >> gpurun_out/summary.txt; }
PT="python -m pytest -q -m gpu --no-header -p no:cacheprovider --tb=short"
run attn      300 $PT tests/test_kernels_gpu.py -k attention -x
run attnbench 200 python tools/attn_bench.py
VGPT_ATTN_V1=1 timeout 200 python tools/attn_bench.py > gpurun_out/attnbench_v1.log 2>&1
for f in attn attnbench attnbench_v1; do echo "=== $f"; tail -n 12 gpurun_out/$f.log | cut -c1-300; done
cat gpurun_out/summary.txt
