#!/bin/bash
# Round 2, call U (1 GPU, ~4 min): the complete -m gpu suite, smoke() and a short default-config bench line on the
# final tree (CFG-branch pairs on peer stores, scheduler update inside every step graph, cost-balanced row shards).
#   gpurun --timeout 420 -- 'bash tools/gpu/r02u.sh'
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
run() { name=$1; shift; local t0=$SECONDS; timeout "$@" > gpurun_out/$name.log 2>&1; local rc=$?; echo "$name exit $rc ($((SECONDS - t0)) s)" >> gpurun_out/summary.txt; return $rc; }
run gpu_suite 300 python -m pytest tests/ -x -q -m gpu --no-header -p no:cacheprovider --tb=short
run smoke 60 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')"
run bench_cfg2 60 python bench.py --steps 3 --warmup 3 --no-baselines --strong none
for f in gpu_suite smoke bench_cfg2; do echo "=== $f"; tail -n ${TAILN:-8} gpurun_out/$f.log 2>/dev/null | cut -c1-600; done
cat gpurun_out/summary.txt
