#!/bin/bash
# Round 2, call M (1 GPU, ~4 min): attention grid with the tail pairs dispatched last; per-round rollout times;
# launch list with the small MLPs on one row (as in the fused sampler loop).
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
run() { name=$1; shift; local t0=$SECONDS; timeout "$@" > gpurun_out/$name.log 2>&1; local rc=$?; echo "$name exit $rc ($((SECONDS - t0)) s)" >> gpurun_out/summary.txt; return $rc; }
PT="python -m pytest -q -m gpu --no-header -p no:cacheprovider --tb=short -x"
run attn_tests 200 $PT tests/test_kernels_gpu.py -k "attention or mask"
run attn_bench 100 python tools/attn_bench.py
run attn_bench_cfg3 100 python tools/attn_bench.py 32 4 256 256
run attn_bench_cfg5 100 python tools/attn_bench.py 4 4 512 512
run sp_tests 300 $PT tests/test_sequence_parallel.py tests/test_zz_batch_gpu.py
run bench_cfg2 200 python bench.py --steps 3 --warmup 3 --no-baselines --strong none
run rollout_bench 300 python tools/rollout_bench.py 8 cfg3
python tools/profile_step.py > gpurun_out/plain.log 2>&1 &&
timeout 300 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/launches.csv python tools/profile_step.py > gpurun_out/ncu_launches.log 2>&1
echo "launches exit $?" >> gpurun_out/summary.txt
for f in attn_tests attn_bench attn_bench_cfg3 attn_bench_cfg5 sp_tests bench_cfg2 rollout_bench; do echo "=== $f"; tail -n ${TAILN:-8} gpurun_out/$f.log 2>/dev/null | cut -c1-400; done
cat gpurun_out/summary.txt
