#!/bin/bash
# Round 2, call L (2 GPUs, ~5 min, charged 2x): the 2-GPU tests with the final kernels and the driver's N = 2 line
# (independent videos + strong-scaling block: cfg5 / cfg3 row-sharded, cfg2 as a CFG-branch pair).
#   gpurun --gpus 2 --timeout 700 -- 'bash tools/gpu/r02l.sh'
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
run() { name=$1; shift; local t0=$SECONDS; timeout "$@" > gpurun_out/$name.log 2>&1; local rc=$?; echo "$name exit $rc ($((SECONDS - t0)) s)" >> gpurun_out/summary.txt; return $rc; }
run mgpu_tests 300 python -m pytest -q -m gpu --no-header -p no:cacheprovider --tb=short tests/test_sequence_parallel.py tests/test_multigpu_gpu.py
run bench_n2 360 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 2 --steps 2 --warmup 3 --no-baselines --strong-timeout 100
for f in mgpu_tests bench_n2; do echo "=== $f"; grep "^{" gpurun_out/$f.log | cut -c1-3500; tail -n 6 gpurun_out/$f.log | grep -v "^{" | cut -c1-300; done
cat gpurun_out/summary.txt
