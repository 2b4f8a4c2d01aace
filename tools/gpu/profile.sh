#!/bin/bash
# parity report + ncu launch list + full captures of the top kernels (1 GPU)
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
run() { name=$1; shift; timeout "$@" > gpurun_out/$name.log 2>&1; echo "$name exit $?" >> gpurun_out/summary.txt; }
run parity 900 python tools/parity_report.py
python tools/profile_step.py > gpurun_out/plain.log 2>&1 &&
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/launches.csv python tools/profile_step.py > gpurun_out/ncu_launches.log 2>&1
echo "launches exit $?" >> gpurun_out/summary.txt
python tools/profile_step.py --no-prefill > gpurun_out/plain2.log 2>&1 &&
ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:gemm_bf16 -s 4 -c 4 \
    -o gpurun_out/prof_gemm -f python tools/profile_step.py --no-prefill > gpurun_out/ncu_gemm.log 2>&1
echo "ncu_gemm exit $?" >> gpurun_out/summary.txt
ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:attn_clip -s 1 -c 1 \
    -o gpurun_out/prof_attn -f python tools/profile_step.py --no-prefill > gpurun_out/ncu_attn.log 2>&1
echo "ncu_attn exit $?" >> gpurun_out/summary.txt
tail -n 12 gpurun_out/parity.log; cat gpurun_out/summary.txt; ls -la gpurun_out
