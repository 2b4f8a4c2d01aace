#!/bin/bash
# Round 2, call D (1 GPU, ~10 min): GEMM with the tail rows in the k-loop (bit-exactness, timing, headline effect), probe
# library, lightweight attention trace, the rewritten full-size parity test, ncu of the new GEMM.
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
run() { name=$1; shift; local t0=$SECONDS; timeout "$@" > gpurun_out/$name.log 2>&1; echo "$name exit $? ($((SECONDS - t0)) s)" >> gpurun_out/summary.txt; }
PT="python -m pytest -q -m gpu --no-header -p no:cacheprovider --tb=short -x"
run tail_tests 300 $PT tests/test_kernels_gpu.py -k "tail or swiglu or gemm"
run kernel_tests 600 $PT tests/test_kernels_gpu.py tests/test_umma_layouts.py
run gemmsweep 120 python tools/gemm_bench.py
run gemmsweep_1040 120 python tools/gemm_bench.py 1040
VGPT_ATTN_VARIANT=8 run attn_trace 100 python tools/attn_trace.py
run bench_cfg2 300 python bench.py --steps 3 --warmup 3 --no-baselines --strong none
VGPT_GEMM_TAIL_IN_LOOP=0 run bench_cfg2_plain 300 python bench.py --steps 3 --warmup 3 --no-baselines --strong none
run model_tests 900 $PT tests/test_model_gpu.py tests/test_zz_batch_gpu.py tests/test_zz_rollout_gpu.py
run fullsize_parity 1500 $PT tests/test_zz_fullsize_gpu.py
python tools/profile_step.py --no-prefill > gpurun_out/plain2.log 2>&1 &&
timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:gemm_bf16 -s 4 -c 4 \
    -o gpurun_out/prof_gemm -f python tools/profile_step.py --no-prefill > gpurun_out/ncu_gemm.log 2>&1
echo "ncu_gemm exit $?" >> gpurun_out/summary.txt
for f in tail_tests kernel_tests gemmsweep gemmsweep_1040 bench_cfg2 bench_cfg2_plain model_tests fullsize_parity; do
  echo "=== $f"; tail -n ${TAILN:-22} gpurun_out/$f.log | cut -c1-400; done
echo "=== attn_trace"; tail -5 gpurun_out/attn_trace.log
cat gpurun_out/summary.txt
