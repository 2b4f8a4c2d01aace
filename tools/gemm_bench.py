#!/usr/bin/env python
"""Time the four projection shapes of the Phi-3 block per tile width (CUDA events, weights of 32
layers cycled so W streams from HBM)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from videogpt_b200 import ops  # noqa: E402

M = int(sys.argv[1]) if len(sys.argv) > 1 else 2064
h, inter, L = 3072, 8192, 16
dev = "cuda"
shapes = {"qkv": (3 * h, h, ops.EPI_STORE), "o": (h, h, ops.EPI_RESIDUAL), "gate_up": (2 * inter, h, ops.EPI_SWIGLU),
          "down": (h, inter, ops.EPI_RESIDUAL)}
for name, (N, K, epi) in shapes.items():
    ws = [(0.02 * torch.randn(N, K, device=dev)).to(torch.bfloat16) for _ in range(L)]
    a = torch.randn(M, K, device=dev).to(torch.bfloat16)
    n_out = N // 2 if epi == ops.EPI_SWIGLU else N
    out = torch.zeros(M, n_out, device=dev, dtype=torch.bfloat16)
    # tail_mode: 1 = plain 256-row tiles, 3 = tail rows in the k-loop of the last full tile row, -1 = library default
    variants = ((256, 1), (192, 1), (256, 3), (192, 3), (0, -1))
    for bn, pair in variants:
        def run():
            for w in ws:
                ops.gemm(a, w, out=out, residual=out if epi == ops.EPI_RESIDUAL else None, epilogue=epi, block_n=bn, tail_mode=pair)
        run(); torch.cuda.synchronize()
        s, t = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(5):
            run()
        t.record(); torch.cuda.synchronize()
        us = s.elapsed_time(t) * 1e3 / (5 * L)
        print(f"{name:8s} M={M} N={N} K={K} block_n={bn} pair={pair}: {us:8.1f} us  {2.0 * M * N * K / us / 1e6:8.1f} TFLOP/s", flush=True)
    del ws
