"""End-to-end parity of the CUDA path (LVM + LVMScheduler of videogpt_b200, through the C ABI)
against the oracle -- the reference's own PyTorch path restated (oracle/), run in bf16 on the
same device with identical weights and latents -- and against the golden outputs of the
unmodified reference (fp32, tests/golden/).

Tolerances are BASELINE.json's: per-step velocity relative L2 <= 1e-2 (bf16) and final-latent
cosine >= 0.999.  The fp32 golden comparison bounds the total bf16 error of the CUDA path and
is reported next to the bf16 oracle's own error against the same golden.
"""
import numpy as np
import pytest
import torch

from oracle import model_oracle as mo, processor_oracle as po, scheduler_oracle as so
from videogpt_b200 import synth

from helpers import FakeTokenizer, FakeVAE, cosine, load_npz, random_pil, rel_l2

pytestmark = pytest.mark.gpu
DEV, BF = "cuda", torch.bfloat16

VEL_TOL = 1e-2      # BASELINE.json: per-step velocity relative L2 error, bf16
COS_TOL = 0.999     # BASELINE.json: final-latent cosine


def _build(dims, seed=0):
    from transformers import Phi3Config
    from videogpt_b200 import LVM
    sd = synth.init_state_dict(dims, seed=seed)
    model = LVM(Phi3Config(**dims.phi3_kwargs()), device=DEV)
    model.load_state_dict(sd)
    model.to(BF).eval()
    return model, sd


def _oracle_cfg(dims):
    return mo.OracleConfig(hidden_size=dims.hidden_size, intermediate_size=dims.intermediate_size,
                           num_hidden_layers=dims.num_hidden_layers, num_attention_heads=dims.num_attention_heads)


def _inputs(n_ctx, n_gen, H, W, device, dtype):
    d = po.frame_block_inputs(n_ctx, n_gen, H, W, True, 1)
    lat = synth.synthetic_latents(n_ctx + n_gen, H, W, seed=42)
    ctx = [x.to(device, dtype) for x in lat[:n_ctx]]
    z0 = [x.to(device, dtype) for x in lat[n_ctx:]]
    mk = dict(input_ids=d["input_ids"].to(device), input_img_latents=ctx,
              input_image_sizes=d["input_image_sizes"], attention_mask=d["attention_mask"].to(device),
              position_ids=d["position_ids"].to(device), denoise_image_sizes=d["denoise_image_sizes"],
              time_emb_inx=d["time_emb_inx"], img_cfg_scale=1.5, use_img_cfg=True, use_kv_cache=False,
              offload_model=False, vae=None)
    return mk, z0


def _oracle_run(sd, dims, mk, z0, steps, pt, dtype, sdpa_backend=None):
    """``sdpa_backend="MATH"``: the same oracle with SDPA forced to the math backend -- a second, independent
    evaluation of the reference's own path in the same precision (the measured bf16-vs-bf16 floor)."""
    import contextlib
    w = {k: v.to(DEV, dtype) for k, v in sd.items()}
    cfg = _oracle_cfg(dims)
    rec = []
    ctx = contextlib.nullcontext()
    if sdpa_backend is not None:
        from torch.nn.attention import SDPBackend, sdpa_kernel
        ctx = sdpa_kernel(getattr(SDPBackend, sdpa_backend))
    with torch.no_grad(), ctx:
        out = so.euler_sample([x.clone() for x in z0] * 2,
                              lambda z, t, **kw: mo.frame_block_forward_with_cfg(w, cfg, z, t, **kw),
                              mk, num_steps=steps, prediction_type=pt, record=rec)
    return out, rec


def _cuda_run(model, mk, z0, steps, pt, use_graph=True):
    from videogpt_b200 import LVMScheduler
    model.use_cuda_graph = use_graph
    sch = LVMScheduler(num_steps=steps)
    sch.record_velocity = []
    out = sch([x.clone() for x in z0] * 2, model.frame_block_forward_with_cfg, mk, use_kv_cache=False,
              prediction_type=pt)
    torch.cuda.synchronize()
    return out, sch.record_velocity


def _gate(floor_fp32, floor_bb):
    """Per-step velocity gate.  BASELINE.json asks for rel-L2 <= 1e-2 against the reference's bf16 path; that is
    asserted as is wherever no amplification applies (step 0, all of v mode at this size -- see the test).  In x1
    mode v = (x1 - z)/(1-sigma) amplifies the bf16 rounding of x1 by 1/(1-sigma), up to 50x at the last step, for
    ANY bf16 evaluation: there the gate is the larger of the MEASURED bf16-vs-bf16 floor (the same oracle with
    SDPA's math backend vs its default backend, `floor_bb`; at full size ours sits within +-2 % of it on all 50
    steps, tests/test_zz_fullsize_gpu.py) and the reference's bf16-vs-fp32 distance (`floor_fp32`: at two layers
    the two SDPA backends differ less than two GEMM accumulation orders do), each with 30 % head-room."""
    return max(VEL_TOL, 1.3 * floor_fp32, 1.3 * floor_bb)


@pytest.mark.parametrize("case", [("tiny", 2, 2, 64, 64, 4), ("ragged", 3, 2, 64, 96, 3), ("cfg1", 4, 4, 256, 256, 4)])
@pytest.mark.parametrize("pt", ["x1", "v"])
def test_next_clip_matches_oracle_bf16(case, pt):
    name, n_ctx, n_gen, H, W, steps = case
    dims = synth.REDUCED
    model, sd = _build(dims)
    mk, z0 = _inputs(n_ctx, n_gen, H, W, DEV, BF)
    want, want_vel = _oracle_run(sd, dims, mk, z0, steps, pt, BF)
    _, alt_vel = _oracle_run(sd, dims, mk, z0, steps, pt, BF, sdpa_backend="MATH")
    mk32, z32 = _inputs(n_ctx, n_gen, H, W, DEV, torch.float32)
    true, true_vel = _oracle_run(sd, dims, mk32, z32, steps, pt, torch.float32)
    got, got_vel = _cuda_run(model, mk, z0, steps, pt)
    assert len(got) == 2 * n_gen and all(torch.equal(got[i], got[n_gen + i]) for i in range(n_gen))   # q7
    for i in range(steps):
        v_ref, v_true = torch.cat(want_vel[i][:n_gen], 0), torch.cat(true_vel[i][:n_gen], 0)
        floor = rel_l2(v_ref, v_true)                       # the reference's own bf16 error
        floor_bb = rel_l2(torch.cat(alt_vel[i][:n_gen], 0), v_ref)      # the reference's bf16 path against itself
        err = rel_l2(got_vel[i], v_ref)
        err_true = rel_l2(got_vel[i], v_true)
        assert err <= _gate(floor, floor_bb), \
            f"step {i}: velocity rel-L2 vs bf16 oracle {err:.3e} (bf16-vs-fp32 {floor:.3e}, bf16-vs-bf16 {floor_bb:.3e})"
        assert err_true <= 1.15 * floor + 1e-3, f"step {i}: vs fp32 oracle {err_true:.3e} (reference bf16: {floor:.3e})"
        if i == 0 or pt == "v":
            assert err <= VEL_TOL, f"step {i}: velocity rel-L2 {err:.3e}"     # no 1/(1-sigma) amplification here
    cos = cosine(torch.cat(got[:n_gen], 0), torch.cat(want[:n_gen], 0))
    assert cos >= COS_TOL, cos


@pytest.mark.parametrize("name,pt", [("tiny_fp32", "x1"), ("tiny_fp32", "v"), ("tiny_ragged_fp32", "x1"),
                                     ("cfg1_fp32", "x1"), ("cfg1_fp32", "v")])
def test_next_clip_vs_reference_golden_fp32(name, pt):
    """CUDA path (bf16) vs the unmodified reference (fp32 CPU golden).  The bound is the bf16
    oracle's own distance to the same golden, times 2."""
    g = load_npz(f"model_{name}.npz")
    n_ctx, n_gen, H, W, steps = [int(x) for x in g["meta"][:5]]
    dims = synth.REDUCED
    model, sd = _build(dims)
    mk, z0 = _inputs(n_ctx, n_gen, H, W, DEV, BF)
    got, _ = _cuda_run(model, mk, z0, steps, pt)
    ref_bf16, _ = _oracle_run(sd, dims, mk, z0, steps, pt, BF)
    gold = torch.from_numpy(g[f"final_{pt}"])[:n_gen].to(DEV)
    e_cuda = rel_l2(torch.cat(got[:n_gen], 0), gold)
    e_ref = rel_l2(torch.cat(ref_bf16[:n_gen], 0), gold)
    assert cosine(torch.cat(got[:n_gen], 0), gold) >= COS_TOL
    assert e_cuda <= max(2.0 * e_ref, 2e-2), (e_cuda, e_ref)


def test_model_callback_seam_matches_engine_loop():
    """Seam S2: driving ``frame_block_forward_with_cfg`` through a foreign scheduler loop (here the
    oracle's) gives the same latents as the fused engine loop, with and without CUDA graphs."""
    dims = synth.REDUCED
    model, sd = _build(dims)
    mk, z0 = _inputs(2, 2, 64, 64, DEV, BF)
    for pt in ("x1", "v"):
        fused, _ = _cuda_run(model, mk, z0, 3, pt, use_graph=True)
        eager, _ = _cuda_run(model, mk, z0, 3, pt, use_graph=False)
        assert all(torch.equal(a, b) for a, b in zip(fused, eager))
        with torch.no_grad():
            seam = so.euler_sample([x.clone() for x in z0] * 2,
                                   lambda z, t, **kw: model.frame_block_forward_with_cfg(z, t, past_key_values=None, **kw)[0],
                                   mk, num_steps=3, prediction_type=pt)
        assert all(torch.equal(a, b) for a, b in zip(fused, seam)), pt


def test_raw_prediction_matches_golden_single_call():
    g = load_npz("model_tiny_fp32.npz")
    dims = synth.REDUCED
    model, sd = _build(dims)
    mk, z0 = _inputs(2, 2, 64, 64, DEV, BF)
    t = torch.full((4,), 0.25, device=DEV)
    pred, cache = model.frame_block_forward_with_cfg([x.clone() for x in z0] * 2, t, past_key_values=None,
                                                     prediction_type="x1", **mk)
    assert cache is None and len(pred) == 4
    gold = torch.from_numpy(g["pred_t025"]).to(DEV)
    assert rel_l2(torch.cat(pred, 0), gold) < 2e-2


def test_context_change_invalidates_prefill():
    dims = synth.REDUCED
    model, sd = _build(dims)
    mk, z0 = _inputs(2, 2, 64, 64, DEV, BF)
    a, _ = _cuda_run(model, mk, z0, 2, "x1")
    mk2 = dict(mk)
    mk2["input_img_latents"] = [x + 0.5 for x in mk["input_img_latents"]]
    b, _ = _cuda_run(model, mk2, z0, 2, "x1")
    c, _ = _cuda_run(model, mk, z0, 2, "x1")
    assert not torch.equal(a[0], b[0]) and torch.equal(a[0], c[0])


def test_wrong_mask_is_rejected():
    dims = synth.REDUCED
    model, sd = _build(dims)
    mk, z0 = _inputs(2, 2, 64, 64, DEV, BF)
    mk["attention_mask"] = torch.ones_like(mk["attention_mask"])
    with pytest.raises(ValueError):
        _cuda_run(model, mk, z0, 1, "x1")
    mk["attention_mask"] = mk["attention_mask"][0]
    with pytest.raises(Exception):
        _cuda_run(model, mk, z0, 1, "x1")


def test_pipeline_next_clip_latents_host_to_host():
    """Public API with HOST latents in and out (what bench.py's e2e leg times)."""
    from videogpt_b200 import LVMPipeline, LVMProcessor
    dims = synth.REDUCED
    model, sd = _build(dims)
    pipe = LVMPipeline(None, model, LVMProcessor(FakeTokenizer()), device=torch.device(DEV))
    lat = synth.synthetic_latents(4, 64, 64, seed=42)
    ctx, noise = lat[:2], lat[2:]
    out = pipe.next_clip_latents([x.pin_memory() for x in ctx], 2, num_inference_steps=4, img_guidance_scale=1.5,
                                 prediction_type="x1", initial_noise=noise)
    mk, z0 = _inputs(2, 2, 64, 64, DEV, BF)
    want, _ = _cuda_run(model, mk, z0, 4, "x1")
    assert len(out) == 2 and all(torch.equal(a, b) for a, b in zip(out, want[:2]))
    seeded = pipe.next_clip_latents(ctx, 2, num_inference_steps=2, img_guidance_scale=1.0, seed=7)
    assert len(seeded) == 2 and all(torch.isfinite(x.float()).all() for x in seeded)


def test_replace_attention_operator_seam():
    """Seam S1: an HF Phi-3 attention module patched by ``replace_attention`` returns what the
    reference's ``new_forward`` computes (oracle ``attention``: qkv, RoPE, dense-mask SDPA, o_proj)."""
    from transformers import Phi3Config
    import transformers.models.phi3.modeling_phi3 as mp
    from videogpt_b200 import replace_attention
    dims = synth.REDUCED
    cfg = Phi3Config(**dims.phi3_kwargs())
    holder = torch.nn.Module()
    holder.attn = mp.Phi3Attention(cfg, layer_idx=0)
    sd = synth.init_state_dict(dims, seed=0, with_pos_embed=False)
    holder.attn.qkv_proj.weight.data.copy_(sd["llm.layers.0.self_attn.qkv_proj.weight"])
    holder.attn.o_proj.weight.data.copy_(sd["llm.layers.0.self_attn.o_proj.weight"])
    holder.to(DEV, BF)
    replace_attention(holder)
    d = po.frame_block_inputs(3, 2, 64, 96, True, 4)
    B, L = d["input_ids"].shape
    g = torch.Generator(device=DEV).manual_seed(5)
    x = torch.randn(B, L, dims.hidden_size, generator=g, device=DEV).to(BF)
    add = mo.additive_mask(d["attention_mask"].to(DEV), BF)
    pos = d["position_ids"].to(DEV)
    out, w, cache = holder.attn(x, attention_mask=add, position_ids=pos, past_key_value=None)
    assert w is None and cache is None and out.shape == x.shape
    ocfg = _oracle_cfg(dims)
    wts = {k: v.to(DEV, BF) for k, v in sd.items() if "layers.0.self_attn" in k}
    cos, sin = mo.rope_cos_sin(pos, ocfg.head_dim, ocfg.rope_theta, BF)
    want = mo.attention(wts, 0, x, add, cos, sin, ocfg)
    assert rel_l2(out, want) < 1e-2                           # all rows, pad rows included
    anti = torch.eye(L, device=DEV).flip(0).bool()            # not of the form code_q >= code_k
    bad = mo.additive_mask(anti[None].expand(B, L, L), BF)
    with pytest.raises(ValueError):
        holder.attn(x, attention_mask=bad, position_ids=pos)


# ---------------------------------------------------------------------------------------------
# the one-frame-at-a-time path (LVM.forward_with_cfg, reference LVM/model.py:330-397, 503-516) and the
# two user entry points of the pipeline (reference LVM/pipeline.py:136-343, 346-595) on the GPU
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("geom", [(2, 64, 64, 1), (3, 48, 80, 4), (4, 256, 256, 1)])
@pytest.mark.parametrize("pt", ["x1", "v"])
def test_single_frame_forward_with_cfg_matches_oracle_bf16(geom, pt):
    """``pipeline.__call__``'s layout: condition tokens (cached prefix) + [time | image] rows, the 1-token
    unconditional row left-padded by the reference.  Same gates as the next-clip path."""
    from videogpt_b200 import LVMScheduler
    n_ctx, H, W, sp = geom
    steps = 4
    dims = synth.REDUCED
    model, sd = _build(dims)
    d = po.single_frame_inputs(n_ctx, H, W, True, sp)
    lat = synth.synthetic_latents(n_ctx + 1, H, W, seed=7)

    def mk_for(dtype):
        return dict(input_ids=d["input_ids"].to(DEV), input_img_latents=[x.to(DEV, dtype) for x in lat[:n_ctx]],
                    input_image_sizes=d["input_image_sizes"], attention_mask=d["attention_mask"].to(DEV),
                    position_ids=d["position_ids"].to(DEV), img_cfg_scale=1.5, use_img_cfg=True, use_kv_cache=False,
                    offload_model=False)

    def oracle(dtype, backend=None):
        import contextlib
        w = {k: v.to(DEV, dtype) for k, v in sd.items()}
        rec = []
        ctx = contextlib.nullcontext()
        if backend is not None:
            from torch.nn.attention import SDPBackend, sdpa_kernel
            ctx = sdpa_kernel(getattr(SDPBackend, backend))
        with torch.no_grad(), ctx:
            out = so.euler_sample(torch.cat([lat[n_ctx]] * 2, 0).to(DEV, dtype),
                                  lambda z, t, **kw: mo.single_frame_forward_with_cfg(w, _oracle_cfg(dims), z, t, **kw),
                                  mk_for(dtype), num_steps=steps, prediction_type=pt, record=rec)
        return out[:1], [r[:1] for r in rec]

    want, want_vel = oracle(BF)
    _, alt_vel = oracle(BF, "MATH")
    true, true_vel = oracle(torch.float32)
    sch = LVMScheduler(num_steps=steps)
    sch.record_velocity = []
    got = sch(torch.cat([lat[n_ctx]] * 2, 0).to(DEV, BF), model.forward_with_cfg, mk_for(BF), use_kv_cache=False,
              prediction_type=pt)
    torch.cuda.synchronize()
    assert torch.equal(got[0], got[1])                                     # cond / uncond duplicates stay identical (q7)
    for i in range(steps):
        floor, floor_bb = rel_l2(want_vel[i], true_vel[i]), rel_l2(alt_vel[i], want_vel[i])
        err, err_true = rel_l2(sch.record_velocity[i], want_vel[i]), rel_l2(sch.record_velocity[i], true_vel[i])
        assert err <= _gate(floor, floor_bb), f"step {i}: {err:.3e} (bf16-vs-fp32 {floor:.3e}, bf16-vs-bf16 {floor_bb:.3e})"
        assert err_true <= 1.15 * floor + 1e-3, f"step {i}: vs fp32 oracle {err_true:.3e} (reference bf16: {floor:.3e})"
        if i == 0 or pt == "v":
            assert err <= VEL_TOL, f"step {i}: velocity rel-L2 {err:.3e}"
    assert cosine(got[:1], want) >= COS_TOL


def _pipeline(model):
    from videogpt_b200 import LVMPipeline, LVMProcessor
    pipe = LVMPipeline(None, model, LVMProcessor(FakeTokenizer()), device=DEV)
    pipe.vae = FakeVAE()
    return pipe


def test_pipeline_frame_block_autoregressive_entry_point_on_gpu():
    """``prompt_condition_frame_block_autoregressive_inference`` (reference pipeline.py:346-595, what the shipped
    inference script calls): PIL frames in, two rounds of two frames out through a fake VAE; every latent handed to
    the decoder against the oracle's frame-block sampler in bf16 on the same context latents and seeded noise."""
    dims = synth.REDUCED
    model, sd = _build(dims)
    pipe = _pipeline(model)
    vae = pipe.vae
    imgs = [random_pil(11), random_pil(12), random_pil(13)]
    out = pipe.prompt_condition_frame_block_autoregressive_inference(
        input_images=imgs, height=64, width=64, gen_nums=[2, 2], num_inference_steps=3, img_guidance_scale=1.5,
        use_input_image_size_as_output=True, dtype=BF, seed=9, prediction_type="x1", clean_image_noise_level=0.0,
        max_frame_window=6)
    assert len(out) == 7 and len(vae.decoded) == 7
    w = {k: v.to(DEV, BF) for k, v in sd.items()}
    frames = list(imgs)
    for k in range(2):
        if k == 1:
            frames = out[:5][-4:]
        ctx = [pipe.vae_encode(pipe.processor.process_image(im).unsqueeze(0).to(DEV), BF) for im in frames]
        d = po.frame_block_inputs(len(ctx), 2, 64, 64, True, 1)
        mk = dict(input_ids=d["input_ids"].to(DEV), input_img_latents=ctx, input_image_sizes=d["input_image_sizes"],
                  attention_mask=d["attention_mask"].to(DEV), position_ids=d["position_ids"].to(DEV),
                  denoise_image_sizes=d["denoise_image_sizes"], time_emb_inx=d["time_emb_inx"], img_cfg_scale=1.5,
                  use_img_cfg=True)
        g = torch.Generator(device=DEV).manual_seed(9)
        noise = [torch.randn(1, 4, 8, 8, device=DEV, generator=g).to(BF) for _ in range(2)]
        with torch.no_grad():
            want = so.euler_sample(noise * 2, lambda z, t, **kw: mo.frame_block_forward_with_cfg(w, _oracle_cfg(dims), z, t, **kw),
                                   mk, num_steps=3, prediction_type="x1")[:2]
        got = [vae.decoded[3 + 2 * k + i].to(DEV) * vae.config.scaling_factor for i in range(2)]
        assert cosine(torch.cat(got, 0), torch.cat(want, 0).float()) >= COS_TOL, k


def test_pipeline_call_one_frame_at_a_time_on_gpu():
    """``LVMPipeline.__call__`` (reference pipeline.py:136-343) on the GPU through a fake VAE."""
    dims = synth.REDUCED
    model, sd = _build(dims)
    pipe = _pipeline(model)
    vae = pipe.vae
    imgs = [random_pil(1), random_pil(2)]
    out = pipe(input_images=imgs, height=64, width=64, gen_num=2, num_inference_steps=3, img_guidance_scale=1.5,
               use_input_image_size_as_output=True, dtype=BF, seed=5, prediction_type="v", clean_image_noise_level=0.0)
    assert len(out) == 4 and len(vae.decoded) == 4
    w = {k: v.to(DEV, BF) for k, v in sd.items()}
    ctx = [pipe.vae_encode(pipe.processor.process_image(im).unsqueeze(0).to(DEV), BF) for im in imgs]
    for k in range(2):
        if k == 1:
            ctx.append(pipe.vae_encode(pipe.processor.process_image(out[2]).unsqueeze(0).to(DEV), BF))
        d = po.single_frame_inputs(len(ctx), 64, 64, True, 1)
        mk = dict(input_ids=d["input_ids"].to(DEV), input_img_latents=ctx, input_image_sizes=d["input_image_sizes"],
                  attention_mask=d["attention_mask"].to(DEV), position_ids=d["position_ids"].to(DEV), img_cfg_scale=1.5,
                  use_img_cfg=True)
        noise = torch.randn(1, 4, 8, 8, device=DEV, generator=torch.Generator(device=DEV).manual_seed(5)).to(BF)
        with torch.no_grad():
            want = so.euler_sample(torch.cat([noise] * 2, 0),
                                   lambda z, t, **kw: mo.single_frame_forward_with_cfg(w, _oracle_cfg(dims), z, t, **kw),
                                   mk, num_steps=3, prediction_type="v")[:1]
        got = vae.decoded[2 + k].to(DEV) * vae.config.scaling_factor
        assert cosine(got, want.float()) >= COS_TOL, k
