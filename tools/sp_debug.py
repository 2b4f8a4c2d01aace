#!/usr/bin/env python
"""Debug driver: unsharded engine vs `world` virtual ranks at full width (few layers) for one
geometry.  python tools/sp_debug.py n_ctx n_gen H W world [layers]"""
import os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import processor_oracle as po
from videogpt_b200 import engine as eng, ops, peer, synth

n_ctx, n_gen, H, W, world = map(int, sys.argv[1:6])
layers = int(sys.argv[6]) if len(sys.argv) > 6 else 2
dev, bf = torch.device("cuda", 0), torch.bfloat16
dims = synth.BackboneDims(num_hidden_layers=layers)
sd = synth.init_state_dict(dims, seed=0, dtype=bf, with_pos_embed=False)
w = eng.EngineWeights(sd, layers, dev)
d = po.frame_block_inputs(n_ctx, n_gen, H, W, True, 1, build_mask=False) if "build_mask" in po.frame_block_inputs.__code__.co_varnames else po.frame_block_inputs(n_ctx, n_gen, H, W, True, 1)
specs, n_lat, n_ctx_lat = eng.frame_block_specs(d["input_ids"], d["position_ids"], d["input_image_sizes"], d["denoise_image_sizes"], d["time_emb_inx"])
lat = synth.synthetic_latents(n_ctx + n_gen, H, W, seed=42)
ctx = torch.cat(lat[:n_ctx], 0).to(dev, bf)
z0 = torch.cat(lat[n_ctx:] * 2, 0).to(dev, bf)
mk = lambda peers=None: eng.NextClipEngine(w, dims.hidden_size, dims.intermediate_size, layers, dims.num_attention_heads, dims.rms_norm_eps, dims.rope_theta, dev, dims.pos_embed_max_size, 2, use_cuda_graph=False, peers=peers)
def sync(tag):
    t0 = time.time(); torch.cuda.synchronize(); print(f"{tag}: ok ({time.time() - t0:.2f}s)", flush=True)
ref = mk()
ref.set_plan(eng.build_plan(specs, n_lat, n_ctx_lat, H // 8, W // 8, dev))
print("plan: prefix rows", ref.plan.prefix.rows, "step rows", ref.plan.step.rows, "pages", ref.plan.total_pages, flush=True)
ref.prefill(ctx); sync("ref prefill")
kv_prefill = ref.kv.clone()
ref.z.copy_(z0); ref.t.fill_(0.3); ref.predict(); sync("ref predict")
ref2 = mk(); ref2.set_plan(eng.build_plan(specs, n_lat, n_ctx_lat, H // 8, W // 8, dev)); ref2.prefill(ctx)
ref2.z.copy_(z0); ref2.t.fill_(0.3); ref2.predict(); sync("ref2")
print("run-to-run deterministic: kv", bool(torch.equal(ref2.kv, ref.kv)), "pred", bool(torch.equal(ref2.pred, ref.pred)), flush=True)
def kvdiff(a, b):
    out = []
    for li in range(a.shape[0]):
        for j, nm in enumerate("kv"):
            d = (a[li, j].float() - b[li, j].float()).abs()
            if d.max() > 0:
                pages = sorted(set(torch.nonzero(d.amax(dim=(1, 2, 3))).flatten().tolist()))
                out.append((li, nm, float(d.max()), pages[:12]))
    return out
if world > 1:
    members = peer.LocalPeerGroup.create(world, dev)
    ranks = [mk(m) for m in members]
    for r, e in enumerate(ranks):
        e.set_plan(eng.build_plan(specs, n_lat, n_ctx_lat, H // 8, W // 8, dev, shard=(r, world)))
        print("rank", r, "prefix", e.plan.prefix.seqs.tolist(), "step", e.plan.step.seqs.tolist(), flush=True)
    eng.run_lockstep([e.prefill_steps(ctx) for e in ranks]); sync("sp prefill")
    print("kv equal after prefill:", [bool(torch.equal(e.kv, kv_prefill)) for e in ranks], flush=True)
    for e in ranks:
        print("  diff:", kvdiff(e.kv, kv_prefill)[:6])
    for e in ranks:
        e.z.copy_(z0); e.t.fill_(0.3)
    eng.run_lockstep([e.predict_steps() for e in ranks]); sync("sp predict")
    print("pred equal:", [bool(torch.equal(e.pred, ref.pred)) for e in ranks], "max abs diff", [float((e.pred.float() - ref.pred.float()).abs().max()) for e in ranks])
    for e in ranks:
        print("  kv diff after predict:", kvdiff(e.kv, ref.kv)[:6])
