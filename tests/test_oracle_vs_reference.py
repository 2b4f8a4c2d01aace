"""Live check of the oracle against the UNMODIFIED reference (build container only;
skipped on the GPU box where /root/reference does not exist)."""
import numpy as np
import pytest
import torch

from oracle import refshim

pytestmark = pytest.mark.skipif(not refshim.reference_available(), reason="reference checkout not present")


def test_processor_live():
    from PIL import Image
    from oracle import processor_oracle as po
    from helpers import prompts
    n_ctx, n_gen, H, W, sp = 3, 2, 64, 96, 4
    proc = refshim.build_reference_processor(sequence_parallel_size=sp)
    imgs = [Image.fromarray(np.zeros((H, W, 3), np.uint8)) for _ in range(n_ctx)]
    p, p_ = prompts(n_ctx, n_gen)
    a = proc.prompt_condition_frame_block_inference([p, p_], [imgs, []], height=H, width=W, use_img_cfg=True,
                                                    use_input_image_size_as_output=True,
                                                    frame_blocks=[n_ctx, n_gen])
    b = po.frame_block_inputs(n_ctx, n_gen, H, W, True, sp)
    for k in ("input_ids", "attention_mask", "position_ids"):
        assert torch.equal(a[k], b[k]), k
    for k in ("input_image_sizes", "denoise_image_sizes", "time_emb_inx", "frame_blocks"):
        assert a[k] == b[k], k


def test_model_and_scheduler_live():
    from LVM.scheduler import LVMScheduler
    from oracle import model_oracle as mo, processor_oracle as po, scheduler_oracle as so
    from videogpt_b200 import synth
    dims = synth.REDUCED
    sd = synth.init_state_dict(dims, seed=3)
    model = refshim.build_reference_model(dims.phi3_kwargs(), sd)
    cfg = mo.OracleConfig(hidden_size=512, intermediate_size=1024, num_hidden_layers=2, num_attention_heads=8)
    n_ctx, n_gen, H, W = 1, 2, 64, 64
    d = po.frame_block_inputs(n_ctx, n_gen, H, W, True, 1)
    lat = synth.synthetic_latents(n_ctx + n_gen, H, W, seed=7)
    ctx, z0 = lat[:n_ctx], lat[n_ctx:]
    mk = dict(input_ids=d["input_ids"], input_img_latents=ctx, input_image_sizes=d["input_image_sizes"],
              attention_mask=d["attention_mask"], position_ids=d["position_ids"],
              denoise_image_sizes=d["denoise_image_sizes"], time_emb_inx=d["time_emb_inx"],
              img_cfg_scale=1.5, use_img_cfg=True, use_kv_cache=False, offload_model=False, vae=None)
    with torch.no_grad():
        for pt in ("x1", "v"):
            ref = LVMScheduler(num_steps=3)([x.clone() for x in z0] * 2, model.frame_block_forward_with_cfg, mk,
                                            use_kv_cache=False, prediction_type=pt)
            got = so.euler_sample([x.clone() for x in z0] * 2,
                                  lambda z, t, **kw: mo.frame_block_forward_with_cfg(sd, cfg, z, t, **kw),
                                  mk, num_steps=3, prediction_type=pt)
            assert all(torch.equal(a, b) for a, b in zip(ref, got))
