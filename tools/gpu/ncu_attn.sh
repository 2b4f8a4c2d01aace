#!/bin/bash
# full ncu capture (with source) of the production attention kernel on the cfg2 step geometry
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
python tools/attn_bench.py > gpurun_out/plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:attn_pair -s 10 -c 1 \
    -o gpurun_out/prof_attn_pair -f python tools/attn_bench.py > gpurun_out/ncu_attn.log 2>&1
echo "ncu exit $?" >> gpurun_out/summary.txt
cat gpurun_out/plain.log; cat gpurun_out/summary.txt
