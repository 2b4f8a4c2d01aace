#!/bin/bash
# Round 2, call J (8 GPUs, ~5 min, charged 8x): the driver's N = 8 line (independent videos + strong-scaling block:
# one cfg5 / cfg3 video row-sharded over all 8 GPUs) and BASELINE configs[3] (32 videos, 4 per rank and pass).
#   gpurun --gpus 8 --timeout 600 -- 'bash tools/gpu/r02j.sh'
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
run() { name=$1; shift; local t0=$SECONDS; timeout "$@" > gpurun_out/$name.log 2>&1; local rc=$?; echo "$name exit $rc ($((SECONDS - t0)) s)" >> gpurun_out/summary.txt; return $rc; }
N=${N:-8}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
nvidia-smi topo -m > gpurun_out/topo.txt 2>&1
run bench_n${N} 330 $TR --master-port 29511 bench.py --gpus $N --steps 2 --warmup 3 --no-baselines --strong-timeout 100
run bench_cfg4_n${N} 200 $TR --master-port 29512 bench.py --gpus $N --config cfg4 --batch 4 --steps 1 --warmup 3 --no-baselines --strong none
for f in bench_n${N} bench_cfg4_n${N}; do echo "=== $f"; grep "^{" gpurun_out/$f.log | cut -c1-3000; tail -n 5 gpurun_out/$f.log | grep -v "^{" | cut -c1-300; done
cat gpurun_out/summary.txt
