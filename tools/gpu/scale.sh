#!/bin/bash
# N-GPU call (N = $1): scaling evidence. dp (independent videos, weak) on cfg2, sp (one video, strong) on cfg5 / cfg3.
N=${1:-8}
mkdir -p gpurun_out
run() { name=$1; shift; timeout "$@" > gpurun_out/$name.log 2>&1; echo "$name exit $?" >> gpurun_out/summary_scale$N.txt; }
rm -f gpurun_out/summary_scale$N.txt
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
run bench_dp${N}_cfg2 240 $TR --master-port 29531 bench.py --gpus $N --steps 3 --warmup 3
run bench_sp${N}_cfg5 300 $TR --master-port 29532 bench.py --gpus $N --steps 2 --warmup 3 --parallelism sp --config cfg5
run bench_sp${N}_cfg3 300 $TR --master-port 29533 bench.py --gpus $N --steps 2 --warmup 3 --parallelism sp --config cfg3
if [ "$N" -ge 4 ]; then run bench_sp2dp_cfg2 240 $TR --master-port 29534 bench.py --gpus $N --steps 3 --warmup 3 --parallelism sp --sp 2; fi
for f in bench_dp${N}_cfg2 bench_sp${N}_cfg5 bench_sp${N}_cfg3 bench_sp2dp_cfg2; do echo "=== $f"; grep "^{" gpurun_out/$f.log 2>/dev/null | cut -c1-330; done
cat gpurun_out/summary_scale$N.txt
