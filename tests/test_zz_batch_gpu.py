"""Batched videos on the GPU (BASELINE.json configs[3]): all rows of several videos and both CFG
branches in one pass.  Every kernel gives a row the same bits whatever else shares its launch
(profiles/r01g_row_partition_invariance.txt), so video v of a batch must equal video v run alone
BIT FOR BIT; the host side of the batched layout is checked against the oracle on CPU
(tests/test_emu_host_numerics.py).  This file sorts last on purpose: it was written after this
round's GPU minutes were spent and runs on hardware for the first time in the round-end suite."""
import pytest
import torch

from videogpt_b200 import synth

pytestmark = pytest.mark.gpu
DEV, BF = "cuda", torch.bfloat16


def _pipe(dims):
    from transformers import Phi3Config
    from videogpt_b200 import LVM, LVMPipeline, LVMProcessor
    sd = synth.init_state_dict(dims, seed=0)
    model = LVM(Phi3Config(**dims.phi3_kwargs()), device=DEV)
    model.load_state_dict(sd)
    model.to(BF).eval()
    return LVMPipeline(None, model, LVMProcessor(synth.SingleIdTagTokenizer()), device=torch.device(DEV))


@pytest.mark.parametrize("pt", ["x1", "v"])
def test_batch_of_videos_equals_the_videos_run_alone(pt):
    n_videos, n_ctx, n_gen, H, W = 3, 2, 2, 128, 128
    pipe = _pipe(synth.REDUCED)
    lats = [synth.synthetic_latents(n_ctx + n_gen, H, W, seed=20 + v) for v in range(n_videos)]
    ctx = [[x.to(DEV, BF) for x in l[:n_ctx]] for l in lats]
    noise = [[x.to(DEV, BF) for x in l[n_ctx:]] for l in lats]
    kw = dict(num_inference_steps=4, img_guidance_scale=1.5, prediction_type=pt)
    got = pipe.next_clip_latents_batch(ctx, n_gen, initial_noise=noise, **kw)
    torch.cuda.synchronize()
    assert len(got) == n_videos
    for v in range(n_videos):
        alone = pipe.next_clip_latents(ctx[v], n_gen, initial_noise=noise[v], **kw)
        for a, b in zip(got[v], alone):
            assert torch.isfinite(a.float()).all()
            assert torch.equal(a, b), f"video {v}: batched and single runs differ"


def test_patch_multiple_resolutions_matches_the_reference_embedders():
    """LVM.patch_multiple_resolutions (reference model.py:292-327) through vgpt_embed_assemble: conv output
    rounded to bf16, then + position embedding rounded to bf16 -- the reference's rounding points, so the
    comparison with the oracle's bf16 PatchEmbedMR is bit-level up to the conv's fp32 summation order."""
    from oracle import model_oracle as mo
    pipe = _pipe(synth.REDUCED)
    m = pipe.model
    sd = {k: v.to(DEV, BF) for k, v in synth.init_state_dict(synth.REDUCED, seed=0).items()}
    cfg = mo.OracleConfig(hidden_size=synth.REDUCED.hidden_size, intermediate_size=synth.REDUCED.intermediate_size,
                          num_hidden_layers=synth.REDUCED.num_hidden_layers,
                          num_attention_heads=synth.REDUCED.num_attention_heads)
    lat = torch.cat(synth.synthetic_latents(3, 64, 96, seed=2), 0).to(DEV, BF)
    for is_ctx, name in ((False, "x_embedder"), (True, "input_x_embedder")):
        got, n_tok, shapes = m.patch_multiple_resolutions(lat, is_input_images=is_ctx)
        want = torch.cat([mo.patch_embed(lat[i:i + 1], sd[f"{name}.proj.weight"], sd[f"{name}.proj.bias"],
                                         sd["pos_embed"], cfg) for i in range(3)], 0)
        assert n_tok == 24 and shapes == [8, 12]
        err = (got.float() - want.float()).abs().max().item()
        assert err <= 2.0 ** -7 * want.float().abs().max().item(), err


def test_submodule_forwards_on_the_gpu():
    """TimestepEmbedder / PatchEmbedMR / FinalLayer forwards of the drop-in (kernel-backed) vs the oracle's
    bf16 modules: within bf16 rounding of the output scale."""
    from oracle import model_oracle as mo
    pipe = _pipe(synth.REDUCED)
    m = pipe.model
    sd = {k: v.to(DEV, BF) for k, v in synth.init_state_dict(synth.REDUCED, seed=0).items()}
    hs = synth.REDUCED.hidden_size

    def close(got, want, tol=2.0 ** -6):
        assert torch.isfinite(got.float()).all()
        assert (got.float() - want.float()).abs().max().item() <= tol * want.float().abs().max().item()

    t = torch.tensor([0.0, 0.25, 0.9], device=DEV)
    close(m.time_token(t), mo.timestep_embedder(sd, "time_token", t, BF))
    close(m.t_embedder(t), mo.timestep_embedder(sd, "t_embedder", t, BF))
    lat = torch.cat(synth.synthetic_latents(2, 64, 96, seed=5), 0).to(DEV, BF)
    want = torch.nn.functional.conv2d(lat, sd["x_embedder.proj.weight"], sd["x_embedder.proj.bias"], stride=2)
    close(m.x_embedder(lat), want.flatten(2).transpose(1, 2))
    g = torch.Generator(device=DEV).manual_seed(1)
    x = torch.randn(2, 24, hs, device=DEV, generator=g).to(BF)
    c = torch.randn(2, hs, device=DEV, generator=g).to(BF)
    close(m.final_layer(x, c), mo.final_layer(sd, x, c), tol=2.0 ** -5)
