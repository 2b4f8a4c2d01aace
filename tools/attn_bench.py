#!/usr/bin/env python
"""Time the attention kernel(s) on a workload geometry (CUDA events), both phases."""
import math
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import processor_oracle as po  # noqa: E402
from videogpt_b200 import engine as eng, ops  # noqa: E402

n_ctx, n_gen, H, W = (int(x) for x in (sys.argv[1:5] if len(sys.argv) > 4 else (4, 4, 256, 256)))
heads, D, dev = (int(sys.argv[5]), int(sys.argv[6]), "cuda") if len(sys.argv) > 6 else (32, 96, "cuda")
d = po.frame_block_inputs(n_ctx, n_gen, H, W, True, 1)
specs, n_lat, n_c = eng.frame_block_specs(d["input_ids"], d["position_ids"], d["input_image_sizes"],
                                          d["denoise_image_sizes"], d["time_emb_inx"])
plan = eng.build_plan(specs, n_lat, n_c, H // 8, W // 8, dev)
L = 8
pools = [(torch.randn(plan.total_pages, heads, 128, D, device=dev).to(torch.bfloat16),
          torch.randn(plan.total_pages, heads, 128, D, device=dev).to(torch.bfloat16)) for _ in range(L)]
bl = H * W // 256 + 2
for phase, ph in (("step", plan.step), ("prefix", plan.prefix)):
    q = torch.randn(ph.rows, 3 * heads * D, device=dev).to(torch.bfloat16)
    out = torch.zeros(ph.rows, heads * D, device=dev, dtype=torch.bfloat16)
    t_ctx, t_gen = n_ctx * bl, n_gen * bl
    flops = 4 * heads * D * (t_gen * (t_ctx + t_gen) + t_gen * t_gen) if phase == "step" else \
        4 * heads * D * bl * bl * sum(range(1, n_ctx + 1))
    for impl in ("tcgen05", "mma_sync"):
        def run():
            for k, v in pools:
                ops.attention(q[:, :heads * D], out, k, v, plan.page_table, ph.seqs, ph.max_q_rows, ph.q_code,
                              plan.k_code, plan.k_tile_minmax, heads, D, 1.0 / math.sqrt(D), impl=impl)
        run(); torch.cuda.synchronize()
        s, t = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(5):
            run()
        t.record(); torch.cuda.synchronize()
        us = s.elapsed_time(t) * 1e3 / (5 * L)
        print(f"{phase:6s} {impl:8s} rows={ph.rows} : {us:8.1f} us  {flops / us / 1e6:8.1f} TFLOP/s (algorithmic)", flush=True)
