#!/bin/bash
# GPU contact script: every stage under its own timeout, logs into gpurun_out/.
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
nproc >> gpurun_out/smi.txt
run() { name=$1; shift; timeout "$@" > gpurun_out/$name.log 2>&1; echo "$name exit $?" >> gpurun_out/summary.txt; }
PT="python -m pytest -q -m gpu --no-header -p no:cacheprovider --tb=short"
run umma      180 $PT tests/test_umma_layouts.py
run gemm      300 $PT tests/test_kernels_gpu.py -k gemm
run kernels   400 $PT tests/test_kernels_gpu.py -k "not gemm"
run model     600 $PT tests/test_model_gpu.py
run smoke     200 python __graft_entry__.py --smoke
run bench1    300 python bench.py --config cfg1 --steps 2 --warmup 3
run bench2    600 python bench.py --steps 2 --warmup 3
for f in umma gemm kernels model smoke bench1 bench2; do echo "=== $f"; tail -n ${TAILN:-25} gpurun_out/$f.log; done
cat gpurun_out/summary.txt
