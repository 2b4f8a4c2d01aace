#!/bin/bash
# Round-2 first contact, part 1 (1 GPU, ~25-30 min of box time): validate what was written after the round-1
# GPU minutes were spent.  Every stage under its own timeout; logs into gpurun_out/.
#   gpurun --timeout 2400 -- 'bash tools/gpu/r02_first.sh'      then tools/gpu/r02_second.sh
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
run() { name=$1; shift; timeout "$@" > gpurun_out/$name.log 2>&1; echo "$name exit $?" >> gpurun_out/summary.txt; }
PT="python -m pytest -q -m gpu --no-header -p no:cacheprovider --tb=short"
# 1. the whole parity suite (includes tests/test_zz_*_gpu.py: batch, rollout, full size -- first hardware run)
run tests        1500 $PT tests
run smoke        200 python __graft_entry__.py --smoke
# 2. opt-in kernels: bit-exactness / tolerance tests, then timings
VGPT_TEST_EXPERIMENTAL=1 run skinny_test 300 $PT tests/test_kernels_gpu.py -k skinny
VGPT_TEST_EXPERIMENTAL=1 run attn_variants 300 $PT tests/test_kernels_gpu.py -k variants
VGPT_TEST_EXPERIMENTAL=1 run fold_test 300 $PT tests/test_model_gpu.py -k folding
for v in 0 1 2 3 4 5; do VGPT_ATTN_VARIANT=$v run attnbench_var$v 200 python tools/attn_bench.py; done
VGPT_ATTN_VARIANT=8 run attn_trace 200 python tools/attn_trace.py
run gemmsweep    300 python tools/gemm_bench.py
VGPT_GEMM_SKINNY_TAIL=1 run gemmsweep_skinny 300 python tools/gemm_bench.py
run umma_rate    200 python tools/umma_rate.py
# 3. the headline line with the default kernels, with the skinny tail, with the attention variants
run bench_cfg2   600 python bench.py --steps 3 --warmup 3
VGPT_GEMM_SKINNY_TAIL=1 run bench_cfg2_skinny 600 python bench.py --steps 3 --warmup 3
VGPT_FOLD_RMSNORM=1 run bench_cfg2_fold 600 python bench.py --steps 3 --warmup 3
VGPT_ATTN_VARIANT=3 run tests_var3 900 $PT tests/test_kernels_gpu.py tests/test_model_gpu.py -k "attention or next_clip"
VGPT_ATTN_VARIANT=3 VGPT_GEMM_SKINNY_TAIL=1 run bench_cfg2_var3_skinny 600 python bench.py --steps 3 --warmup 3
for f in tests smoke skinny_test attn_variants fold_test bench_cfg2_fold attnbench_var0 attnbench_var1 attnbench_var2 attnbench_var3 attnbench_var4 \
         attnbench_var5 tests_var3 bench_cfg2 bench_cfg2_skinny bench_cfg2_var3_skinny; do
  echo "=== $f"; tail -n ${TAILN:-5} gpurun_out/$f.log | cut -c1-500; done
cat gpurun_out/summary.txt
