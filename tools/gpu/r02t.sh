#!/bin/bash
# Round 2, call T (2 GPUs, ~2 min, charged 2x): after call S found ~70 ms per clip of exposed host time in every peer
# group (the reference's end-of-sampler gc.collect, run while the GPU idles behind the clip-boundary synchronisation):
# the CFG-branch pair and the row-sharded cfg5 video again, without it and with the cost-balanced row partition
# (cheapest-with-dearest chunk pairs, remainder rows of the two sequences on different ranks).
#   gpurun --gpus 2 --timeout 330 -- 'bash tools/gpu/r02t.sh'
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
run() { name=$1; shift; local t0=$SECONDS; timeout "$@" > gpurun_out/$name.log 2>&1; local rc=$?; echo "$name exit $rc ($((SECONDS - t0)) s)" >> gpurun_out/summary.txt; return $rc; }
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
run mgpu_tests 120 python -m pytest -q -m gpu --no-header -p no:cacheprovider --tb=short tests/test_sequence_parallel.py -k "two_gpus and (geom2 or geom1)"
run bench_cfgpair 100 $TR --master-port 29531 bench.py --gpus 2 --parallelism cfg --config cfg2 --steps 3 --warmup 3 --no-baselines --strong none
run bench_sp2_cfg5 120 $TR --master-port 29532 bench.py --gpus 2 --parallelism sp --config cfg5 --steps 2 --warmup 3 --no-baselines --strong none
for f in mgpu_tests bench_cfgpair bench_sp2_cfg5; do echo "=== $f"; grep "^{" gpurun_out/$f.log | cut -c1-3800; tail -n 4 gpurun_out/$f.log | grep -v "^{" | cut -c1-300; done
cat gpurun_out/summary.txt
