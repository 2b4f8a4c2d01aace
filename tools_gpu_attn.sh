#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
run() { name=$1; shift; timeout "$@" > gpurun_out/$name.log 2>&1; echo "$name exit $?" >> gpurun_out/summary.txt; }
PT="python -m pytest -q -m gpu --no-header -p no:cacheprovider --tb=short"
run umma      200 $PT tests/test_umma_layouts.py
run attn      400 $PT tests/test_kernels_gpu.py -k attention
run attnbench 300 python tools/attn_bench.py
run model     900 $PT tests/test_model_gpu.py
run bench2    600 python bench.py --steps 3 --warmup 3
for f in umma attn attnbench model bench2; do echo "=== $f"; tail -n 12 gpurun_out/$f.log | cut -c1-400; done
cat gpurun_out/summary.txt
