#!/bin/bash
# Round 2, call K (1 GPU, ~8 min): the complete -m gpu suite as the driver runs it (final kernels), rollout lines
# (persistent K/V cache vs the reference's recompute flow), --set full captures of the elementwise kernels.
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
run() { name=$1; shift; local t0=$SECONDS; timeout "$@" > gpurun_out/$name.log 2>&1; local rc=$?; echo "$name exit $rc ($((SECONDS - t0)) s)" >> gpurun_out/summary.txt; return $rc; }
run gpu_suite 900 python -m pytest tests/ -x -q -m gpu --no-header -p no:cacheprovider --tb=short
run attn_bench 100 python tools/attn_bench.py
run bench_rollout 300 python bench.py --config cfg3 --rollout 4 --steps 1 --warmup 1 --no-baselines --strong none
run bench_rollout_recompute 300 python bench.py --config cfg3 --rollout 4 --recompute --steps 1 --warmup 1 --no-baselines --strong none
timeout 300 ncu --profile-from-start off --set full --clock-control none --import-source on \
    -k "regex:rmsnorm|rope_kv|linear_small|embed_assemble|timestep" -c 10 \
    -o gpurun_out/prof_small -f python tools/profile_step.py --no-prefill > gpurun_out/ncu_small.log 2>&1
echo "ncu_small exit $?" >> gpurun_out/summary.txt
timeout 300 ncu --profile-from-start off --set full --clock-control none --import-source on \
    -k "regex:final_layer|cfg_euler" -c 2 \
    -o gpurun_out/prof_final -f python tools/profile_step.py --no-prefill > gpurun_out/ncu_final.log 2>&1
echo "ncu_final exit $?" >> gpurun_out/summary.txt
for f in gpu_suite attn_bench bench_rollout bench_rollout_recompute; do echo "=== $f"; tail -n ${TAILN:-8} gpurun_out/$f.log 2>/dev/null | cut -c1-400; done
cat gpurun_out/summary.txt
