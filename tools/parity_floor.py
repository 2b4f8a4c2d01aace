#!/usr/bin/env python
"""What "two bf16 evaluations of the reference path" agree to, measured -- and where the CUDA path stands
against it, per Euler step over the full 50-step clip and per decoder layer.

BASELINE.json states the tolerance as: per-step velocity rel-L2 <= 1e-2 (bf16) against the reference's own
PyTorch path, final-latent cosine >= 0.999.  Round 1 found ours-vs-oracle-bf16 above 1e-2 at full size and
argued from oracle-bf16 vs oracle-fp32 that no two bf16 evaluations can meet it.  This tool measures that
claim directly: the SAME oracle (the reference path restated, as written: no cache, padded unconditional
row, dense mask) in bf16, evaluated independently

  A  on the GPU, eager, default SDPA backend (cuBLAS + fused attention)        <- the gate's reference
  B  on the GPU, SDPA forced to the MATH backend (bmm / softmax / bmm in bf16)
  C  on the host CPU (oneDNN bf16 GEMMs, CPU SDPA)                              (first --cpu-steps steps)
  F  on the GPU in fp32                                                         <- ground truth

every one following its OWN trajectory from the same noise, like the CUDA path ("ours") does.  Reported per
step: ours|B|C vs A (the gate and its floor), and ours|A|B vs F (absolute error).  Second part: hidden
state after every decoder layer of the first forward (t = 0), ours vs F and A vs F, generated rows and
context rows separately, so that a kernel regression cannot hide inside the end-to-end floor.

    python tools/parity_floor.py --out gpurun_out/parity_floor.json [--steps 50] [--reduced]
"""
import argparse
import contextlib
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import model_oracle as mo, processor_oracle as po, scheduler_oracle as so  # noqa: E402
from videogpt_b200 import LVM, LVMScheduler, synth  # noqa: E402

DEV, BF = "cuda", torch.bfloat16


def sync():
    if DEV != "cpu":
        torch.cuda.synchronize()


def rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def cos(a, b):
    a, b = a.double().cpu().flatten(), b.double().cpu().flatten()
    return float(a @ b / (a.norm() * b.norm()))


class Case:
    def __init__(self, dims, n_ctx, n_gen, H, W):
        from transformers import Phi3Config
        self.dims, self.n_ctx, self.n_gen, self.H, self.W = dims, n_ctx, n_gen, H, W
        self.sd = synth.init_state_dict(dims, seed=0, device=DEV, with_pos_embed=False)
        self.sd["pos_embed"] = synth.sincos_pos_embed_table(dims.hidden_size, dims.pos_embed_max_size).to(DEV)
        self.model = LVM(Phi3Config(**dims.phi3_kwargs()), device=DEV, materialize_pos_embed=False)
        self.model.load_state_dict(self.sd, strict=False)
        self.model.pos_embed = self.sd["pos_embed"]
        self.model.to(BF).eval()
        self.d = po.frame_block_inputs(n_ctx, n_gen, H, W, True, 1)
        self.lat = synth.synthetic_latents(n_ctx + n_gen, H, W, seed=42)
        self.cfg = mo.OracleConfig(hidden_size=dims.hidden_size, intermediate_size=dims.intermediate_size,
                                   num_hidden_layers=dims.num_hidden_layers, num_attention_heads=dims.num_attention_heads)
        self._w = {}

    def mk(self, dtype, dev):
        d = self.d
        return dict(input_ids=d["input_ids"].to(dev), input_img_latents=[x.to(dev, dtype) for x in self.lat[:self.n_ctx]],
                    input_image_sizes=d["input_image_sizes"], attention_mask=d["attention_mask"].to(dev),
                    position_ids=d["position_ids"].to(dev), denoise_image_sizes=d["denoise_image_sizes"],
                    time_emb_inx=d["time_emb_inx"], img_cfg_scale=1.5, use_img_cfg=True, use_kv_cache=False,
                    offload_model=False, vae=None)

    def weights(self, dtype, dev):
        key = (dtype, dev)
        if key not in self._w:
            self._w = {k: v for k, v in self._w.items() if k[0] == BF and k[1] == DEV}    # keep at most bf16-GPU + one more
            self._w[key] = {k: v.to(dev, dtype) for k, v in self.sd.items()}
        return self._w[key]

    def oracle(self, dtype, steps, pt, dev=None, backend=None):
        dev = DEV if dev is None else dev
        w = self.weights(dtype, dev)
        rec = []
        ctx = contextlib.nullcontext()
        if backend is not None:
            from torch.nn.attention import sdpa_kernel, SDPBackend
            ctx = sdpa_kernel(getattr(SDPBackend, backend))
        t0 = time.perf_counter()
        with torch.no_grad(), ctx:
            out = so.euler_sample([x.to(dev, dtype) for x in self.lat[self.n_ctx:]] * 2,
                                  lambda z, t, **kw: mo.frame_block_forward_with_cfg(w, self.cfg, z, t, **kw),
                                  self.mk(dtype, dev), num_steps=steps, prediction_type=pt, record=rec)
        if dev == DEV:
            sync()
        n = self.n_gen
        return torch.cat(out[:n], 0).cpu(), [torch.cat(r[:n], 0).cpu() for r in rec], time.perf_counter() - t0

    def ours(self, steps, pt):
        sch = LVMScheduler(num_steps=steps)
        sch.record_velocity = []
        out = sch([x.to(DEV, BF) for x in self.lat[self.n_ctx:]] * 2, self.model.frame_block_forward_with_cfg,
                  self.mk(BF, DEV), use_kv_cache=False, prediction_type=pt)
        sync()
        return torch.cat(out[:self.n_gen], 0).cpu(), [v.cpu() for v in sch.record_velocity]


def trajectories(case, steps, pt, cpu_steps):
    """Every evaluation follows its own trajectory; the CPU one only for its first `cpu_steps` steps (the
    sigma grid of a shorter run differs, so C runs the full-length grid and is cut by an exception)."""
    res, t = {}, {}
    res["ours"] = case.ours(steps, pt)
    fa, va, t["A"] = case.oracle(BF, steps, pt)
    fb, vb, t["B"] = case.oracle(BF, steps, pt, backend="MATH")
    ff, vf, t["F"] = case.oracle(torch.float32, steps, pt)
    res.update(A=(fa, va), B=(fb, vb), F=(ff, vf))
    vc = []
    if cpu_steps > 0:
        class _Stop(Exception):
            pass
        w = case.weights(BF, "cpu")
        torch.set_num_threads(os.cpu_count() or 1)

        def f(z, tt, **kw):
            if len(vc_raw) >= cpu_steps:
                raise _Stop
            return mo.frame_block_forward_with_cfg(w, case.cfg, z, tt, **kw)
        vc_raw = []
        t0 = time.perf_counter()
        try:
            with torch.no_grad():
                so.euler_sample([x.to(BF) for x in case.lat[case.n_ctx:]] * 2, f, case.mk(BF, "cpu"), num_steps=steps,
                                prediction_type=pt, record=vc_raw)
        except _Stop:
            pass
        t["C"] = time.perf_counter() - t0
        vc = [torch.cat(r[:case.n_gen], 0) for r in vc_raw]
        case._w.pop((BF, "cpu"), None)
    fo, vo = res["ours"]
    out = {"prediction_type": pt, "steps": steps, "seconds": t,
           "what": {"A": "oracle bf16, GPU eager, default SDPA", "B": "oracle bf16, GPU, SDPA MATH backend",
                    "C": f"oracle bf16, host CPU ({os.cpu_count()} threads), first {cpu_steps} steps",
                    "F": "oracle fp32, GPU", "ours": "videogpt_b200 CUDA path (fused loop, CUDA graph)"},
           "vel_ours_vs_A": [rel(a, b) for a, b in zip(vo, va)],
           "vel_B_vs_A": [rel(a, b) for a, b in zip(vb, va)],
           "vel_C_vs_A": [rel(a, b) for a, b in zip(vc, va)],
           "vel_ours_vs_F": [rel(a, b) for a, b in zip(vo, vf)],
           "vel_A_vs_F": [rel(a, b) for a, b in zip(va, vf)],
           "vel_B_vs_F": [rel(a, b) for a, b in zip(vb, vf)],
           "final_cos": {"ours_vs_A": cos(fo, fa), "B_vs_A": cos(fb, fa), "ours_vs_F": cos(fo, ff), "A_vs_F": cos(fa, ff),
                         "B_vs_F": cos(fb, ff)},
           "final_rel": {"ours_vs_A": rel(fo, fa), "B_vs_A": rel(fb, fa), "ours_vs_F": rel(fo, ff), "A_vs_F": rel(fa, ff),
                         "B_vs_F": rel(fb, ff)}}
    over = lambda xs: sum(x > 1e-2 for x in xs)
    out["steps_over_1e-2"] = {"ours_vs_A": over(out["vel_ours_vs_A"]), "B_vs_A": over(out["vel_B_vs_A"]),
                              "A_vs_F": over(out["vel_A_vs_F"])}
    # is the CUDA path inside the bf16-vs-bf16 spread, step by step?
    out["max_ratio_ours_over_B"] = max(a / max(b, 1e-30) for a, b in zip(out["vel_ours_vs_A"], out["vel_B_vs_A"]))
    out["max_ratio_oursF_over_AF"] = max(a / max(b, 1e-30) for a, b in zip(out["vel_ours_vs_F"], out["vel_A_vs_F"]))
    return out


def per_layer(case):
    """Hidden state after every decoder layer of the first forward (t = 0): ours (eager, tapped) and oracle
    bf16 (A) against oracle fp32 (F); generated rows of both CFG rows, and context rows (ours: the prefill)."""
    m, d = case.model, case.d
    n_ctx, n_gen = case.n_ctx, case.n_gen
    L = d["input_ids"].shape[1]
    bl = (L // (n_ctx + n_gen))
    t_ctx, t_gen = n_ctx * bl, n_gen * bl
    z = [x for x in case.lat[n_ctx:]] * 2
    ts = torch.zeros(len(z))

    def orc(dtype):
        w = case.weights(dtype, DEV)
        mk = case.mk(dtype, DEV)
        layers = []
        with torch.no_grad():
            mo.frame_block_forward(w, case.cfg, [x.to(DEV, dtype) for x in z], ts.to(DEV), mk["input_ids"], mk["input_img_latents"],
                                   mk["input_image_sizes"], mk["attention_mask"], mk["position_ids"], mk["denoise_image_sizes"],
                                   mk["time_emb_inx"], layer_outputs=layers)
        gen = [torch.cat([h[0, t_ctx:], h[1, L - t_gen:]], 0).cpu() for h in layers]
        ctx = [h[0, :t_ctx].cpu() for h in layers]
        return gen, ctx

    gen_a, ctx_a = orc(BF)
    gen_f, ctx_f = orc(torch.float32)
    m.use_cuda_graph = False
    m._engine = None
    mk = case.mk(BF, DEV)
    e = m.engine()
    e.layer_tap = []
    m.frame_block_forward([x.to(DEV, BF) for x in z], ts.to(DEV), mk["input_ids"], mk["input_img_latents"], mk["input_image_sizes"],
                          mk["attention_mask"], mk["position_ids"], mk["denoise_image_sizes"], mk["time_emb_inx"])
    sync()
    taps = e.layer_tap
    e.layer_tap = None
    m.use_cuda_graph = True
    m._engine = None
    nl = case.cfg.num_hidden_layers
    # prefill runs layers 0..L-2 completely (the last one stops after its K/V append), then the step runs all L
    ctx_o = [t.cpu() for t in taps[:len(taps) - nl]]
    gen_o = [t.cpu() for t in taps[len(taps) - nl:]]
    out = {"rows": {"generated": 2 * t_gen, "context": t_ctx},
           "gen_ours_vs_F": [rel(a, b) for a, b in zip(gen_o, gen_f)],
           "gen_A_vs_F": [rel(a, b) for a, b in zip(gen_a, gen_f)],
           "gen_ours_vs_A": [rel(a, b) for a, b in zip(gen_o, gen_a)],
           "ctx_ours_vs_F": [rel(a, b) for a, b in zip(ctx_o, ctx_f)],
           "ctx_A_vs_F": [rel(a, b) for a, b in zip(ctx_a, ctx_f)]}
    out["max_ratio_gen_ours_over_A"] = max(a / max(b, 1e-30) for a, b in zip(out["gen_ours_vs_F"], out["gen_A_vs_F"]))
    out["max_ratio_ctx_ours_over_A"] = max(a / max(b, 1e-30) for a, b in zip(out["ctx_ours_vs_F"], out["ctx_A_vs_F"]))
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "parity_floor.json"))
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--cpu-steps", type=int, default=2)
    ap.add_argument("--reduced", action="store_true", help="reduced backbone, small frames (a quick check of the tool)")
    ap.add_argument("--modes", default="x1,v")
    args = ap.parse_args()
    if args.reduced:
        case, name = Case(synth.REDUCED, 2, 2, 64, 64), "reduced_2+2x64x64"
    else:
        case, name = Case(synth.FULL_SIZE, 4, 4, 256, 256), "full_cfg2"
    report = {"case": name, "gpu": torch.cuda.get_device_name(0) if DEV != "cpu" else "cpu", "torch": torch.__version__}
    report["per_layer_t0"] = per_layer(case)
    pl = report["per_layer_t0"]
    print("per layer (t=0): gen rows ours/F last %.3e  A/F last %.3e  max ratio ours:A %.3f | ctx rows max ratio %.3f" % (
        pl["gen_ours_vs_F"][-1], pl["gen_A_vs_F"][-1], pl["max_ratio_gen_ours_over_A"], pl["max_ratio_ctx_ours_over_A"]), flush=True)
    for pt in args.modes.split(","):
        r = trajectories(case, args.steps, pt, args.cpu_steps if pt == "x1" else 0)
        report[f"{name}_{args.steps}steps_{pt}"] = r
        print(f"{pt}: steps over 1e-2 {r['steps_over_1e-2']}  max vel ours/A %.3e  B/A %.3e  C/A %s  A/F %.3e  ours/F %.3e | "
              "cos ours/A %.6f B/A %.6f A/F %.6f ours/F %.6f | max ratio ours:B %.2f, oursF:AF %.3f | s %s" % (
                  max(r["vel_ours_vs_A"]), max(r["vel_B_vs_A"]), ["%.3e" % x for x in r["vel_C_vs_A"]], max(r["vel_A_vs_F"]),
                  max(r["vel_ours_vs_F"]), r["final_cos"]["ours_vs_A"], r["final_cos"]["B_vs_A"], r["final_cos"]["A_vs_F"],
                  r["final_cos"]["ours_vs_F"], r["max_ratio_ours_over_B"], r["max_ratio_oursF_over_AF"],
                  {k: round(v, 1) for k, v in r["seconds"].items()}), flush=True)
        os.makedirs(os.path.dirname(args.out), exist_ok=True)
        with open(args.out, "w") as f:
            json.dump(report, f, indent=1)
    with open(args.out, "w") as f:
        json.dump(report, f, indent=1)


if __name__ == "__main__":
    main()
