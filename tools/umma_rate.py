#!/usr/bin/env python
"""Cycles per tcgen05.mma (M=128, cta_group::1) by shape / operand form (vgpt_debug_umma_rate)."""
import ctypes, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from videogpt_b200 import _lib
names = {0: "SS K-major SW128", 1: "SS K-major SW64", 2: "TS + MN-major SW128 B", 3: "TS + MN-major SW64 B"}
ctas = 148
out = torch.zeros(ctas, device="cuda")
for mode in (0, 1, 2, 3):
    for n_acc, ce in ((1, 0), (2, 0), (1, 8), (1, 4), (1, 1)):
        row = []
        for N in (32, 64, 96, 128, 192, 256):
            if n_acc == 2 and N > 128:
                continue
            _lib.call_probe("vgpt_debug_umma_rate", mode, N, 4096, n_acc, ce, ctas, ctypes.c_void_p(out.data_ptr()), torch.cuda.current_stream().cuda_stream)
            torch.cuda.synchronize()
            row.append(f"N={N}: {out.mean().item():6.1f}")
        print(f"{names[mode]:24s} acc={n_acc} commit_every={ce}  " + "  ".join(row) + "   (floor N/2)", flush=True)

# CTA pairs (the GEMM's instruction): 74 clusters
out2 = torch.zeros(74, device="cuda")
row = []
for N in (16, 32, 48, 64, 128, 192, 224, 256):
    _lib.call_probe("vgpt_debug_umma_rate", 4, N, 4096, 1, 0, 74, ctypes.c_void_p(out2.data_ptr()), torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    row.append(f"N={N}: {out2.mean().item():6.1f}")
print("SS cta_group::2 (M=256)      " + "  ".join(row) + "   (floor N/2)", flush=True)
