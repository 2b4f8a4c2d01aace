#!/bin/bash
# Round 2, call S (2 GPUs, ~2.5 min, charged 2x): CFG-branch pairs on the peer-store mechanism (partition="sequences":
# prediction pushed to the peer, two barriers per step, update inside the step graph) and the sequence-parallel step
# with the update inside the graph; the NCCL all-gather form runs once more beside it as the cross-check.
#   gpurun --gpus 2 --timeout 420 -- 'bash tools/gpu/r02s.sh'
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
run() { name=$1; shift; local t0=$SECONDS; timeout "$@" > gpurun_out/$name.log 2>&1; local rc=$?; echo "$name exit $rc ($((SECONDS - t0)) s)" >> gpurun_out/summary.txt; return $rc; }
run mgpu_tests 200 python -m pytest -q -m gpu --no-header -p no:cacheprovider --tb=short tests/test_sequence_parallel.py tests/test_multigpu_gpu.py
run bench_cfgpair 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 2 --parallelism cfg --config cfg2 --steps 3 --warmup 3 --no-baselines --strong none
for f in mgpu_tests bench_cfgpair; do echo "=== $f"; grep "^{" gpurun_out/$f.log | cut -c1-3500; tail -n 8 gpurun_out/$f.log | grep -v "^{" | cut -c1-300; done
cat gpurun_out/summary.txt
