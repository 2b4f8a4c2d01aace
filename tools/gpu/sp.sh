#!/bin/bash
# SP validation: 1-GPU part (virtual ranks, invariance) always; 2-GPU part when 2 devices are visible
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
run() { name=$1; shift; timeout "$@" > gpurun_out/$name.log 2>&1; echo "$name exit $?" >> gpurun_out/summary.txt; }
PT="python -m pytest -q -m gpu --no-header -p no:cacheprovider --tb=short"
run sp_tests 600 $PT tests/test_sequence_parallel.py tests/test_multigpu_gpu.py tests/test_kernels_gpu.py -k "sequence or virtual or cfg_branch or attention"
run invariance 200 python tools/invariance_probe.py
if [ "$(nvidia-smi -L | wc -l)" -ge 2 ]; then
  run bench_sp2_cfg2 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 2 --steps 3 --warmup 3 --parallelism sp
fi
for f in sp_tests invariance bench_sp2_cfg2; do echo "=== $f"; grep -v "Warn\|gemm " gpurun_out/$f.log 2>/dev/null | tail -n 14 | cut -c1-1200; done
cat gpurun_out/summary.txt
