"""Generate the committed golden fixtures by running the UNMODIFIED reference on CPU.

Run in the build container only (needs /root/reference):

    PYTHONPATH=/root/repo python tests/golden/make_golden.py

Writes ``tests/golden/processor_golden.json`` (hashes of the reference's masks /
positions / index dicts) and ``tests/golden/model_*.npz`` (outputs of the
reference's ``LVMScheduler`` driving the reference's ``LVM`` on deterministic
synthetic weights and latents, fp32).  The reference has no golden vectors of its
own, so these are the pins for ``oracle/`` and, through it,
for the CUDA path.
"""
from __future__ import annotations

import hashlib
import json
import os
import sys

import numpy as np
import torch
from PIL import Image

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle import refshim  # noqa: E402
from videogpt_b200 import synth  # noqa: E402

PROCESSOR_CASES = [  # (n_ctx, n_gen, H, W, sp)
    (4, 4, 256, 256, 1), (4, 4, 512, 512, 1), (32, 4, 256, 256, 1), (8, 4, 176, 320, 4),
    (3, 5, 176, 320, 8), (3, 4, 176, 320, 8), (5, 2, 64, 96, 8), (1, 1, 64, 64, 1), (2, 3, 64, 64, 1),
]


def h16(t: torch.Tensor) -> str:
    return hashlib.sha256(t.contiguous().numpy().tobytes()).hexdigest()[:16]


def prompts(n_ctx, n_gen):
    p = "".join(f"<img><|image_{i + 1}|></img>" for i in range(n_ctx))
    p += "".join(f"<|diffusion|><|image_{n_ctx + i + 1}|>" for i in range(n_gen))
    p_ = "".join(f"<|diffusion|><|image_{i + 1}|>" for i in range(n_gen))
    return p, p_


def processor_goldens():
    out = []
    for n_ctx, n_gen, H, W, sp in PROCESSOR_CASES:
        proc = refshim.build_reference_processor(sequence_parallel_size=sp)
        imgs = [Image.fromarray(np.zeros((H, W, 3), np.uint8)) for _ in range(n_ctx)]
        p, p_ = prompts(n_ctx, n_gen)
        d = proc.prompt_condition_frame_block_inference(
            [p, p_], [imgs, []], height=H, width=W, use_img_cfg=True,
            use_input_image_size_as_output=True, frame_blocks=[n_ctx, n_gen])
        m, pos = d["attention_mask"], d["position_ids"]
        entry = {
            "case": [n_ctx, n_gen, H, W, sp], "path": "frame_block", "L": int(m.shape[-1]),
            "mask_dtype": str(m.dtype), "mask_sha": h16(m), "mask_ones": int(m.sum()),
            "pos_sha": h16(pos), "pos_sum": int(pos.sum()), "ids_sha": h16(d["input_ids"]),
            "input_image_sizes": {str(k): v for k, v in d["input_image_sizes"].items()},
            "denoise_image_sizes": {str(k): v for k, v in d["denoise_image_sizes"].items()},
            "time_emb_inx": {str(k): v for k, v in d["time_emb_inx"].items()},
            "row_sums_first3_cond": [[int(x) for x in m[0, s - 1:s + 2].sum(-1)]
                                     for s, _ in (d["input_image_sizes"][0] + [[r[0] - 1, 0] for r in d["denoise_image_sizes"][0]])],
        }
        out.append(entry)
        if n_ctx * H * W <= 8 * 176 * 320:   # the one-frame-at-a-time path, smaller cases only
            d = proc([p.split("<|diffusion|>")[0]], [imgs], height=H, width=W, use_img_cfg=True,
                     use_input_image_size_as_output=True)
            m, pos = d["attention_mask"], d["position_ids"]
            out.append({"case": [n_ctx, 1, H, W, sp], "path": "single_frame", "L": int(m.shape[-1]),
                        "mask_dtype": str(m.dtype), "mask_sha": h16(m), "mask_ones": int(m.sum()),
                        "pos_sha": h16(pos), "pos_sum": int(pos.sum()), "ids_sha": h16(d["input_ids"]),
                        "input_image_sizes": {str(k): v for k, v in d["input_image_sizes"].items()}})
    return out


def model_kwargs_from(d, ctx_latents, guidance=1.5):
    return dict(input_ids=d["input_ids"], input_img_latents=ctx_latents,
                input_image_sizes=d["input_image_sizes"], attention_mask=d["attention_mask"],
                position_ids=d["position_ids"], denoise_image_sizes=d["denoise_image_sizes"],
                time_emb_inx=d["time_emb_inx"], img_cfg_scale=guidance, use_img_cfg=True,
                use_kv_cache=False, offload_model=False, vae=None)


def model_golden(name, dims, n_ctx, n_gen, H, W, steps):
    from LVM.scheduler import LVMScheduler
    sd = synth.init_state_dict(dims, seed=0)
    model = refshim.build_reference_model(dims.phi3_kwargs(), sd)
    proc = refshim.build_reference_processor()
    imgs = [Image.fromarray(np.zeros((H, W, 3), np.uint8)) for _ in range(n_ctx)]
    p, p_ = prompts(n_ctx, n_gen)
    d = proc.prompt_condition_frame_block_inference(
        [p, p_], [imgs, []], height=H, width=W, use_img_cfg=True,
        use_input_image_size_as_output=True, frame_blocks=[n_ctx, n_gen])
    lat = synth.synthetic_latents(n_ctx + n_gen, H, W, seed=42)
    ctx, z0 = lat[:n_ctx], lat[n_ctx:]
    mk = model_kwargs_from(d, ctx)
    arrays = {"meta": np.array([n_ctx, n_gen, H, W, steps, dims.hidden_size, dims.intermediate_size,
                                dims.num_hidden_layers, dims.num_attention_heads])}
    with torch.no_grad():
        # one raw model call at t = 0.25 (both branches, uncombined x1 predictions)
        z = [x.clone() for x in z0] * 2
        t = torch.full((len(z),), 0.25)
        pred, _ = model.frame_block_forward_with_cfg(z, t, past_key_values=None, prediction_type="x1", **mk)
        arrays["pred_t025"] = torch.cat(pred, 0).numpy()
        for pt in ("x1", "v"):
            z = [x.clone() for x in z0] * 2
            out = LVMScheduler(num_steps=steps)(z, model.frame_block_forward_with_cfg, mk,
                                                use_kv_cache=False, prediction_type=pt)
            arrays[f"final_{pt}"] = torch.cat(out, 0).numpy()
    np.savez_compressed(os.path.join(HERE, f"model_{name}.npz"), **arrays)
    print(name, {k: v.shape for k, v in arrays.items()})


def single_frame_golden(name, dims, n_ctx, H, W, steps):
    """``pipeline.__call__`` path: ``LVM.forward_with_cfg`` (needs a process group: model.py:371)."""
    import torch.distributed as dist
    from LVM.scheduler import LVMScheduler
    if not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29517")
        dist.init_process_group("gloo", rank=0, world_size=1)
    sd = synth.init_state_dict(dims, seed=0)
    model = refshim.build_reference_model(dims.phi3_kwargs(), sd)
    proc = refshim.build_reference_processor()
    imgs = [Image.fromarray(np.zeros((H, W, 3), np.uint8)) for _ in range(n_ctx)]
    p = "".join(f"<img><|image_{i + 1}|></img>" for i in range(n_ctx))
    d = proc([p], [imgs], height=H, width=W, use_img_cfg=True, use_input_image_size_as_output=True)
    lat = synth.synthetic_latents(n_ctx + 1, H, W, seed=42)
    ctx, z0 = lat[:n_ctx], lat[n_ctx]
    mk = dict(input_ids=d["input_ids"], input_img_latents=ctx, input_image_sizes=d["input_image_sizes"],
              attention_mask=d["attention_mask"], position_ids=d["position_ids"], img_cfg_scale=1.5,
              use_img_cfg=True, use_kv_cache=False, offload_model=False)
    arrays = {"meta": np.array([n_ctx, 1, H, W, steps, dims.hidden_size, dims.intermediate_size,
                                dims.num_hidden_layers, dims.num_attention_heads])}
    with torch.no_grad():
        for pt in ("x1", "v"):
            z = torch.cat([z0, z0], 0)
            out = LVMScheduler(num_steps=steps)(z, model.forward_with_cfg, mk, use_kv_cache=False,
                                                prediction_type=pt)
            arrays[f"final_{pt}"] = out.numpy()
    np.savez_compressed(os.path.join(HERE, f"model_{name}.npz"), **arrays)
    print(name, {k: v.shape for k, v in arrays.items()})


if __name__ == "__main__":
    torch.manual_seed(0)
    with open(os.path.join(HERE, "processor_golden.json"), "w") as f:
        json.dump(processor_goldens(), f, indent=1)
    model_golden("tiny_fp32", synth.REDUCED, 2, 2, 64, 64, 4)
    model_golden("tiny_ragged_fp32", synth.REDUCED, 3, 2, 64, 96, 3)
    model_golden("cfg1_fp32", synth.REDUCED, 4, 4, 256, 256, 4)      # BASELINE.json configs[0]
    single_frame_golden("single_frame_fp32", synth.REDUCED, 2, 64, 64, 3)
