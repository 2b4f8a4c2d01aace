#!/bin/bash
# 2-GPU call: CFG-split parity, DP and CFG-split bench lines
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
run() { name=$1; shift; timeout "$@" > gpurun_out/$name.log 2>&1; echo "$name exit $?" >> gpurun_out/summary.txt; }
PT="python -m pytest -q -m gpu --no-header -p no:cacheprovider --tb=short"
nvidia-smi -L > gpurun_out/gpus.txt
run multigpu 600 $PT tests/test_multigpu_gpu.py
run bench_dp2 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3
run bench_cfg2 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 3 --warmup 3 --parallelism cfg
for f in multigpu bench_dp2 bench_cfg2; do echo "=== $f"; tail -n 8 gpurun_out/$f.log | cut -c1-700; done
cat gpurun_out/summary.txt
