#!/bin/bash
# GPU contact script: every stage under its own timeout, logs into gpurun_out/.
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
run() { name=$1; shift; timeout "$@" > gpurun_out/$name.log 2>&1; echo "$name exit $?" >> gpurun_out/summary.txt; }
PT="python -m pytest -q -m gpu --no-header -p no:cacheprovider --tb=short"
run tests     1500 $PT tests
run gemmsweep 300 python tools/gemm_bench.py
run smoke     200 python __graft_entry__.py --smoke
run bench2    600 python bench.py --steps 3 --warmup 3
python tools/profile_step.py > gpurun_out/plain.log 2>&1 &&
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/launches.csv python tools/profile_step.py > gpurun_out/ncu_launches.log 2>&1
echo "launches exit $?" >> gpurun_out/summary.txt
for f in tests gemmsweep smoke bench2; do echo "=== $f"; tail -n ${TAILN:-25} gpurun_out/$f.log; done
cat gpurun_out/summary.txt
