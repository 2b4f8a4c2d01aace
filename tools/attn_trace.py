#!/usr/bin/env python
"""Timeline of ONE attention CTA (grid block (0, 0, 0): head 0, first query pair of the first sequence):

    VGPT_ATTN_VARIANT=8 python tools/attn_trace.py [n_ctx n_gen H W]      (needs a B200)

The trace variant of attn_pair_tcgen05_kernel records a clock64 stamp at every hand-over between
its roles; this prints them per KV tile, in cycles relative to the CTA's start: when the MMA warp
could issue S / P V (and how long it waited for the softmax warps), when each softmax warpgroup
received S, finished its exponentials, got the shared P buffer and published P, and when the TMA
warp re-filled a K/V stage.  One run answers "which unit waits for which" -- the question the
83 us / 53 us / 42 us gap of DESIGN.md section 4.2 leaves open."""
import ctypes
import math
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import processor_oracle as po  # noqa: E402
from videogpt_b200 import _lib, engine as eng, ops  # noqa: E402

EV = {1: "mma: S may issue", 2: "mma: S issued", 3: "mma: P ready", 4: "mma: PV issued", 5: "mma: loop top", 6: "mma: fenced", 10: "sm: S full", 11: "sm: S read",
      12: "sm: exp done", 13: "sm: P buffer free", 14: "sm: P written", 15: "sm: epilogue", 20: "tma: stage free",
      21: "tma: stage issued", 30: "start", 31: "table done", 32: "end"}


def main():
    if os.environ.get("VGPT_ATTN_VARIANT") != "8":
        sys.exit("run with VGPT_ATTN_VARIANT=8")
    n_ctx, n_gen, H, W = (int(x) for x in (sys.argv[1:5] if len(sys.argv) > 4 else (4, 4, 256, 256)))
    heads, D, dev = 32, 96, "cuda"
    d = po.frame_block_inputs(n_ctx, n_gen, H, W, True, 1)
    specs, n_lat, n_c = eng.frame_block_specs(d["input_ids"], d["position_ids"], d["input_image_sizes"],
                                              d["denoise_image_sizes"], d["time_emb_inx"])
    plan = eng.build_plan(specs, n_lat, n_c, H // 8, W // 8, dev)
    ph = plan.step
    k = torch.randn(plan.total_pages, heads, 128, D, device=dev).to(torch.bfloat16)
    v = torch.randn(plan.total_pages, heads, 128, D, device=dev).to(torch.bfloat16)
    q = torch.randn(ph.rows, 3 * heads * D, device=dev).to(torch.bfloat16)
    out = torch.zeros(ph.rows, heads * D, device=dev, dtype=torch.bfloat16)
    buf = np.zeros(2 * 8192, np.uint64)
    n = ctypes.c_int(0)
    stream = torch.cuda.current_stream().cuda_stream

    def launch():
        ops.attention(q[:, :heads * D], out, k, v, plan.page_table, ph.seqs, ph.max_q_rows, ph.q_code, plan.k_code,
                      plan.k_tile_minmax, heads, D, 1.0 / math.sqrt(D))

    launch()                                                            # warm-up launch, its trace is discarded
    _lib.call("vgpt_debug_attn_trace", buf.ctypes.data_as(ctypes.c_void_p), 8192, ctypes.byref(n), stream)
    launch()
    _lib.call("vgpt_debug_attn_trace", buf.ctypes.data_as(ctypes.c_void_p), 8192, ctypes.byref(n), stream)
    ev = sorted((int(buf[2 * i]), int(buf[2 * i + 1])) for i in range(n.value))
    t0 = ev[0][0]
    print(f"{n.value} events; cycles relative to the CTA's start")
    last = {}
    for clk, tag in ev:
        warp, tile, j, e = tag >> 40, (tag >> 32) & 0xff, (tag >> 8) & 0xffffff, tag & 0xff
        who = "AB"[tile] if e < 20 else "-"
        key = (warp, tile)
        print(f"{clk - t0:9d}  (+{clk - last.get(key, clk):6d} in this role)  warp {warp:2d} tile {who}  kv {j:3d}  {EV.get(e, e)}")
        last[key] = clk
    # summary: time per KV tile seen by the MMA warp, and its two waits
    per = {}
    for clk, tag in ev:
        e, j, tile = tag & 0xff, (tag >> 8) & 0xffffff, (tag >> 32) & 0xff
        per.setdefault((j, tile), {})[e] = clk - t0
    pv = sorted((j, x[4]) for (j, t), x in per.items() if t == 0 and 4 in x)
    if len(pv) > 2:
        gaps = [b[1] - a[1] for a, b in zip(pv, pv[1:])]
        print(f"\nP V_A issue period: median {sorted(gaps)[len(gaps) // 2]} cycles per KV tile (tensor-pipe work of a pair of tiles: 2212)")


if __name__ == "__main__":
    main()
