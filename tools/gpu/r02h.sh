#!/bin/bash
# Round 2, call H (1 GPU, ~10 min): attention with FMA-pipe masking on mixed tiles + finer MMA-warp trace; split-K gone;
# the default bench line as the driver runs it; 1-GPU lines of cfg3 / cfg5 / cfg4 (sample of 8 videos); launch list and
# --set full captures of the attention kernel and the four GEMMs of a step.
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
run() { name=$1; shift; local t0=$SECONDS; timeout "$@" > gpurun_out/$name.log 2>&1; local rc=$?; echo "$name exit $rc ($((SECONDS - t0)) s)" >> gpurun_out/summary.txt; return $rc; }
PT="python -m pytest -q -m gpu --no-header -p no:cacheprovider --tb=short -x"
run kernel_tests 300 $PT tests/test_kernels_gpu.py tests/test_umma_layouts.py
run attn_bench 100 python tools/attn_bench.py
run attn_bench_cfg3 100 python tools/attn_bench.py 32 4 256 256
run attn_bench_cfg5 100 python tools/attn_bench.py 4 4 512 512
VGPT_ATTN_VARIANT=8 run attn_trace 100 python tools/attn_trace.py
run model_tests 600 $PT tests/test_model_gpu.py tests/test_zz_batch_gpu.py tests/test_zz_rollout_gpu.py tests/test_sequence_parallel.py
run bench_default 500 python bench.py
run bench_cfg3 200 python bench.py --config cfg3 --steps 2 --warmup 3 --no-baselines --strong none
run bench_cfg5 200 python bench.py --config cfg5 --steps 2 --warmup 3 --no-baselines --strong none
run bench_cfg4_b4 300 python bench.py --config cfg4 --batch 4 --videos 8 --steps 1 --warmup 3 --no-baselines --strong none
run smoke 200 python __graft_entry__.py --smoke
python tools/profile_step.py > gpurun_out/plain.log 2>&1 &&
timeout 300 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/launches.csv python tools/profile_step.py > gpurun_out/ncu_launches.log 2>&1
echo "launches exit $?" >> gpurun_out/summary.txt
timeout 300 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:attn_pair -s 40 -c 2 \
    -o gpurun_out/prof_attn -f python tools/profile_step.py > gpurun_out/ncu_attn.log 2>&1
echo "ncu_attn exit $?" >> gpurun_out/summary.txt
timeout 300 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:gemm_bf16 -s 4 -c 4 \
    -o gpurun_out/prof_gemm -f python tools/profile_step.py --no-prefill > gpurun_out/ncu_gemm.log 2>&1
echo "ncu_gemm exit $?" >> gpurun_out/summary.txt
for f in kernel_tests attn_bench attn_bench_cfg3 attn_bench_cfg5 model_tests bench_default bench_cfg3 bench_cfg5 bench_cfg4_b4 smoke; do
  echo "=== $f"; tail -n ${TAILN:-12} gpurun_out/$f.log 2>/dev/null | cut -c1-330; done
cat gpurun_out/summary.txt
