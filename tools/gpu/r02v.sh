#!/bin/bash
# Round 2, call V (1 GPU, ~2 min): `python bench.py` exactly as the driver runs it, on the final tree.
#   gpurun --timeout 190 -- 'bash tools/gpu/r02v.sh'
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
t0=$SECONDS; timeout 170 python bench.py > gpurun_out/bench_default.log 2> gpurun_out/bench_default.err; echo "bench_default exit $? ($((SECONDS - t0)) s)" >> gpurun_out/summary.txt
grep "^{" gpurun_out/bench_default.log | cut -c1-6000; tail -n 5 gpurun_out/bench_default.err | cut -c1-300
cat gpurun_out/summary.txt
