#!/usr/bin/env python
"""Per-step parity numbers of the CUDA path against the oracle (the reference's PyTorch path
restated) on the same device: velocity rel-L2 per Euler step and final-latent cosine, for

  ours (bf16 kernels)  vs  oracle bf16      -- the BASELINE.json gate
  ours                 vs  oracle fp32      -- absolute error of the CUDA path
  oracle bf16          vs  oracle fp32      -- the reference's own bf16 noise floor

Writes one JSON document (default gpurun_out/parity.json; copy under profiles/ to commit)."""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import model_oracle as mo, processor_oracle as po, scheduler_oracle as so  # noqa: E402
from videogpt_b200 import LVM, LVMScheduler, synth  # noqa: E402

DEV, BF = "cuda", torch.bfloat16


def rel(a, b):
    a, b = a.double(), b.double()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def cos(a, b):
    a, b = a.double().flatten(), b.double().flatten()
    return float(a @ b / (a.norm() * b.norm()))


def run(dims, n_ctx, n_gen, H, W, steps, pt, seed=0):
    from transformers import Phi3Config
    sd = synth.init_state_dict(dims, seed=seed, device=DEV, with_pos_embed=False)
    sd["pos_embed"] = synth.sincos_pos_embed_table(dims.hidden_size, dims.pos_embed_max_size).to(DEV)
    model = LVM(Phi3Config(**dims.phi3_kwargs()), device=DEV, materialize_pos_embed=False)
    model.load_state_dict(sd, strict=False)
    model.pos_embed = sd["pos_embed"]
    model.to(BF).eval()
    d = po.frame_block_inputs(n_ctx, n_gen, H, W, True, 1)
    lat = synth.synthetic_latents(n_ctx + n_gen, H, W, seed=42)

    def mk_for(dtype):
        return dict(input_ids=d["input_ids"].to(DEV), input_img_latents=[x.to(DEV, dtype) for x in lat[:n_ctx]],
                    input_image_sizes=d["input_image_sizes"], attention_mask=d["attention_mask"].to(DEV),
                    position_ids=d["position_ids"].to(DEV), denoise_image_sizes=d["denoise_image_sizes"],
                    time_emb_inx=d["time_emb_inx"], img_cfg_scale=1.5, use_img_cfg=True, use_kv_cache=False,
                    offload_model=False, vae=None)

    cfg = mo.OracleConfig(hidden_size=dims.hidden_size, intermediate_size=dims.intermediate_size,
                          num_hidden_layers=dims.num_hidden_layers, num_attention_heads=dims.num_attention_heads)

    def oracle(dtype):
        w = {k: v.to(dtype) for k, v in sd.items()}
        rec = []
        with torch.no_grad():
            out = so.euler_sample([x.to(DEV, dtype) for x in lat[n_ctx:]] * 2,
                                  lambda z, t, **kw: mo.frame_block_forward_with_cfg(w, cfg, z, t, **kw),
                                  mk_for(dtype), num_steps=steps, prediction_type=pt, record=rec)
        return torch.cat(out[:n_gen], 0), [torch.cat(r[:n_gen], 0) for r in rec]

    sch = LVMScheduler(num_steps=steps)
    sch.record_velocity = []
    ours = sch([x.to(DEV, BF) for x in lat[n_ctx:]] * 2, model.frame_block_forward_with_cfg, mk_for(BF),
               use_kv_cache=False, prediction_type=pt)
    ours_final, ours_vel = torch.cat(ours[:n_gen], 0), sch.record_velocity
    ref16_final, ref16_vel = oracle(BF)
    ref32_final, ref32_vel = oracle(torch.float32)
    del model
    torch.cuda.empty_cache()
    return {
        "prediction_type": pt, "steps": steps,
        "velocity_rel_l2_ours_vs_oracle_bf16": [rel(a, b) for a, b in zip(ours_vel, ref16_vel)],
        "velocity_rel_l2_ours_vs_oracle_fp32": [rel(a, b) for a, b in zip(ours_vel, ref32_vel)],
        "velocity_rel_l2_oracle_bf16_vs_fp32": [rel(a, b) for a, b in zip(ref16_vel, ref32_vel)],
        "final_cosine_ours_vs_oracle_bf16": cos(ours_final, ref16_final),
        "final_cosine_ours_vs_oracle_fp32": cos(ours_final, ref32_final),
        "final_cosine_oracle_bf16_vs_fp32": cos(ref16_final, ref32_final),
        "final_rel_l2_ours_vs_oracle_bf16": rel(ours_final, ref16_final),
        "final_rel_l2_ours_vs_oracle_fp32": rel(ours_final, ref32_final),
        "final_rel_l2_oracle_bf16_vs_fp32": rel(ref16_final, ref32_final),
    }


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "parity.json"))
    ap.add_argument("--full-steps", type=int, default=10)
    ap.add_argument("--skip-full", action="store_true")
    args = ap.parse_args()
    report = {}
    for pt in ("x1", "v"):
        report[f"reduced_cfg1_{pt}"] = run(synth.REDUCED, 4, 4, 256, 256, 4, pt)
        report[f"reduced_50steps_{pt}"] = run(synth.REDUCED, 2, 2, 64, 64, 50, pt)
    if not args.skip_full:
        for pt in ("x1", "v"):
            report[f"full_cfg2_{args.full_steps}steps_{pt}"] = run(synth.FULL_SIZE, 4, 4, 256, 256, args.full_steps, pt)
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    with open(args.out, "w") as f:
        json.dump(report, f, indent=1)
    for k, v in report.items():
        print(k, "max vel err vs bf16 oracle %.3e | ours vs fp32 %.3e | oracle bf16 vs fp32 %.3e | cos %.6f" % (
            max(v["velocity_rel_l2_ours_vs_oracle_bf16"]), max(v["velocity_rel_l2_ours_vs_oracle_fp32"]),
            max(v["velocity_rel_l2_oracle_bf16_vs_fp32"]), v["final_cosine_ours_vs_oracle_bf16"]))


if __name__ == "__main__":
    main()
