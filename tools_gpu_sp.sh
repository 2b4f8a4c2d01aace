mkdir -p gpurun_out
PT="python -m pytest -q -m gpu --no-header -p no:cacheprovider --tb=short"
timeout 600 $PT tests/test_sequence_parallel.py tests/test_multigpu_gpu.py -x > gpurun_out/sp.log 2>&1; echo "sp exit $?"; tail -30 gpurun_out/sp.log
