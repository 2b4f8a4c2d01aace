#!/bin/bash
# Round 2, call B (1 GPU, ~8 min): fused tail tiles of the GEMM (bit-exactness, timing, effect on the headline),
# attention trace + variants after the setmaxnreg fix.
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
run() { name=$1; shift; local t0=$SECONDS; timeout "$@" > gpurun_out/$name.log 2>&1; echo "$name exit $? ($((SECONDS - t0)) s)" >> gpurun_out/summary.txt; }
PT="python -m pytest -q -m gpu --no-header -p no:cacheprovider --tb=short -x"
run tail_tests 300 $PT tests/test_kernels_gpu.py -k "fused_tail"
run kernel_tests 600 $PT tests/test_kernels_gpu.py
run gemmsweep 120 python tools/gemm_bench.py
run gemmsweep_1040 120 python tools/gemm_bench.py 1040
VGPT_ATTN_VARIANT=8 run attn_trace 100 python tools/attn_trace.py
VGPT_TEST_EXPERIMENTAL=1 run attn_variants 200 $PT tests/test_kernels_gpu.py -k variants
for v in 0 1 2 3 4 5; do VGPT_ATTN_VARIANT=$v run attnbench_var$v 60 python tools/attn_bench.py; done
run model_tests 600 $PT tests/test_model_gpu.py
run bench_cfg2 300 python bench.py --steps 2 --warmup 3 --no-baselines --strong none
VGPT_GEMM_FUSED_TAIL=0 run bench_cfg2_notail 300 python bench.py --steps 2 --warmup 3 --no-baselines --strong none
run bench_cfg2_strong1 400 python bench.py --steps 2 --warmup 3 --no-baselines
for f in tail_tests kernel_tests gemmsweep gemmsweep_1040 attn_variants attnbench_var0 attnbench_var1 attnbench_var2 attnbench_var3 attnbench_var4 attnbench_var5 model_tests bench_cfg2 bench_cfg2_notail bench_cfg2_strong1; do
  echo "=== $f"; tail -n ${TAILN:-6} gpurun_out/$f.log | cut -c1-600; done
echo "=== attn_trace"; head -60 gpurun_out/attn_trace.log
cat gpurun_out/summary.txt
