#!/bin/bash
# Round 2, call W (1 GPU, < 1 min): what the 16 tag / time rows of cfg2 (M = 2064 = 8 x 256 + 16) cost each projection:
# the GEMM sweep at M = 2048 next to M = 2064.
#   gpurun --timeout 100 -- 'bash tools/gpu/r02w.sh'
mkdir -p gpurun_out
timeout 45 python tools/gemm_bench.py 2048 > gpurun_out/gemmsweep_2048.log 2>&1; echo "gemmsweep_2048 exit $?"
timeout 45 python tools/gemm_bench.py 2064 > gpurun_out/gemmsweep_2064.log 2>&1; echo "gemmsweep_2064 exit $?"
paste -d'\n' gpurun_out/gemmsweep_2048.log gpurun_out/gemmsweep_2064.log | cut -c1-120
