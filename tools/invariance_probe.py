#!/usr/bin/env python
"""Is every kernel's per-row result independent of how rows are grouped into launches / tiles?
(the property sequence parallelism relies on for bit-exactness)"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import processor_oracle as po
from videogpt_b200 import engine as eng, ops, synth

dev, bf = torch.device("cuda", 0), torch.bfloat16
torch.manual_seed(0)
M, h, I = 2064, 3072, 8192
for name, N, K, epi in (("qkv", 3 * h, h, ops.EPI_STORE), ("o", h, h, ops.EPI_RESIDUAL), ("gate_up", 2 * I, h, ops.EPI_SWIGLU), ("down", h, I, ops.EPI_RESIDUAL)):
    a = torch.randn(M, K, device=dev).to(bf)
    w = (torch.randn(N, K, device=dev) * 0.02).to(bf)
    r = torch.randn(M, N, device=dev).to(bf) if epi == ops.EPI_RESIDUAL else None
    full = ops.gemm(a, w, residual=r, epilogue=epi)
    for lo, hi in ((0, 1032), (1032, 2064), (516, 1032), (100, 358)):
        part = ops.gemm(a[lo:hi], w, residual=None if r is None else r[lo:hi].contiguous(), epilogue=epi)
        same = torch.equal(part, full[lo:hi])
        print(f"gemm {name:8s} rows [{lo},{hi}): bit-equal {same}" + ("" if same else f"  max diff {float((part.float() - full[lo:hi].float()).abs().max()):.4g}, rows differing {int((part != full[lo:hi]).any(1).sum())}"), flush=True)
    for bn in (128, 192, 256):
        alt = ops.gemm(a, w, residual=r, epilogue=epi, block_n=bn, tail_mode=1)
        print(f"gemm {name:8s} block_n {bn}: bit-equal {torch.equal(alt, full)}", flush=True)

# attention: cfg2 geometry, sharded q rows vs all rows
H, D = 32, 96
d = po.frame_block_inputs(4, 4, 256, 256, True, 1)
specs, n_lat, n_ctx_lat = eng.frame_block_specs(d["input_ids"], d["position_ids"], d["input_image_sizes"], d["denoise_image_sizes"], d["time_emb_inx"])
full = eng.build_plan(specs, n_lat, n_ctx_lat, 32, 32, dev)
kp = torch.randn(full.total_pages, H, 128, D, device=dev).to(bf)
vp = torch.randn(full.total_pages, H, 128, D, device=dev).to(bf)
for impl in ("tcgen05", "mma_sync"):
  for which in ("prefix", "step"):
    ph = getattr(full, which)
    q = torch.randn(ph.rows, H * D, device=dev).to(bf)
    out = torch.zeros(ph.rows, H * D, device=dev, dtype=bf)
    ops.attention(q, out, kp, vp, full.page_table, ph.seqs, ph.max_q_rows, ph.q_code, full.k_code, full.k_tile_minmax, H, D, D ** -0.5, impl=impl)
    out2 = torch.zeros_like(out)
    ops.attention(q, out2, kp, vp, full.page_table, ph.seqs, ph.max_q_rows, ph.q_code, full.k_code, full.k_tile_minmax, H, D, D ** -0.5, impl=impl)
    print(impl, which, "run-to-run equal", torch.equal(out, out2))
    for world in (2,):
        for r in range(world):
            p = eng.build_plan(specs, n_lat, n_ctx_lat, 32, 32, dev, shard=(r, world))
            sph = getattr(p, which)
            idx = []
            for s_, sp in enumerate(specs):
                lo, hi = (0, sp.n_prefix) if which == "prefix" else (sp.n_prefix, sp.n_prefix + sp.n_active)
                base = int(ph.seqs[s_, 0]) - lo
                for a_, b_ in eng.shard_ranges(lo, hi, r, world, flip=bool(s_ & 1)):     # as build_plan deals them
                    idx.append(torch.arange(a_ + base, b_ + base, device=dev))
            idx = torch.cat(idx)
            ql = q[idx].contiguous()
            ol = torch.zeros(len(idx), H * D, device=dev, dtype=bf)
            ops.attention(ql, ol, kp, vp, p.page_table, sph.seqs, sph.max_q_rows, sph.q_code, p.k_code, p.k_tile_minmax, H, D, D ** -0.5, impl=impl)
            ne = (ol != out[idx])
            rows = torch.nonzero(ne.any(1)).flatten().tolist()
            print(f"{impl} attention {which} world {world} rank {r}: rows differing {len(rows)}/{len(idx)}", flush=True)
            for lr in rows[:8]:
                gr = int(idx[lr])
                heads = sorted(set((torch.nonzero(ne[lr]).flatten() // D).tolist()))
                dmax = float((ol[lr].float() - out[gr].float()).abs().max())
                print(f"   local row {lr} (tile {lr // 128} off {lr % 128}) global row {gr} (tile {gr // 128} off {gr % 128}) code {int(ph.q_code[gr])}: heads {heads[:10]} n_heads {len(heads)} cols differing {int(ne[lr].sum())} max abs {dmax:.4g}")
