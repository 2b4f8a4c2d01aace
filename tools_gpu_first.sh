#!/bin/bash
# GPU contact script: every stage under its own timeout, logs into gpurun_out/.
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
run() { name=$1; shift; timeout "$@" > gpurun_out/$name.log 2>&1; echo "$name exit $?" >> gpurun_out/summary.txt; }
PT="python -m pytest -q -m gpu --no-header -p no:cacheprovider --tb=short"
run umma      180 $PT tests/test_umma_layouts.py
run attn      400 $PT tests/test_kernels_gpu.py -k attention
run kernels   400 $PT tests/test_kernels_gpu.py -k "not attention"
run model     900 $PT tests/test_model_gpu.py
run gemmsweep 300 python tools/gemm_bench.py
run smoke     200 python __graft_entry__.py --smoke
run bench2    600 python bench.py --steps 3 --warmup 3
for f in umma attn kernels model gemmsweep smoke bench2; do echo "=== $f"; tail -n ${TAILN:-25} gpurun_out/$f.log; done
cat gpurun_out/summary.txt
