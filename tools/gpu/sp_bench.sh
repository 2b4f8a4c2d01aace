#!/bin/bash
# N-GPU call (N = $1, default 2): sequence-parallel bench lines for cfg2 / cfg5 / cfg3
N=${1:-2}
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
run() { name=$1; shift; timeout "$@" > gpurun_out/$name.log 2>&1; echo "$name exit $?" >> gpurun_out/summary.txt; }
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
run bench_sp${N}_cfg2 600 $TR --master-port 29521 bench.py --gpus $N --steps 3 --warmup 3 --parallelism sp
run bench_sp${N}_cfg5 900 $TR --master-port 29522 bench.py --gpus $N --steps 2 --warmup 3 --parallelism sp --config cfg5
run bench_sp${N}_cfg3 900 $TR --master-port 29523 bench.py --gpus $N --steps 2 --warmup 3 --parallelism sp --config cfg3
for f in bench_sp${N}_cfg2 bench_sp${N}_cfg5 bench_sp${N}_cfg3; do echo "=== $f"; grep -v Warning gpurun_out/$f.log | tail -n 4 | cut -c1-900; done
cat gpurun_out/summary.txt
