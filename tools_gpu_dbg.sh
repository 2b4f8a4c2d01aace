timeout 200 python -m pytest -q -m gpu --no-header -p no:cacheprovider --tb=short tests/test_kernels_gpu.py -k attention -x 2>&1 | tail -3
for f in 0 1; do echo "== flags $f"; VGPT_DEBUG_ATTN_FLAGS=$f timeout 60 python tools/attn_bench.py 2>&1 | grep tcgen05; done
echo "== cfg3 geometry (73 kv tiles)"; for f in 0; do VGPT_DEBUG_ATTN_FLAGS=$f timeout 60 python tools/attn_bench.py 32 4 256 256 2>&1 | grep "tcgen05"; done
echo "== cfg5 geometry"; timeout 60 python tools/attn_bench.py 4 4 512 512 2>&1 | grep "tcgen05"
