"""Numerical check of the HOST side on CPU: the product's model / scheduler / engine run with the
torch emulation of the kernels (tests/emu_ops.py, same paged-cache / code / row-array layouts) in
fp32 and must reproduce the oracle -- the reference's path as written: no cache, padded
unconditional row, dense mask -- to fp32 rounding.  A wrong K/V slot, token code, RoPE position,
latent number or cached-prefix decision fails here without a GPU; the kernels' own arithmetic is
the business of the -m gpu tests."""
import pytest
import torch

from oracle import model_oracle as mo, processor_oracle as po, scheduler_oracle as so
from videogpt_b200 import synth

from helpers import FakeVAE as _FakeVAE, random_pil as _pil

TOL = 2e-5


def _model(dims=synth.REDUCED, seed=0):
    from transformers import Phi3Config
    from videogpt_b200 import LVM
    sd = synth.init_state_dict(dims, seed=seed)
    m = LVM(Phi3Config(**dims.phi3_kwargs()), device="cpu")
    m.load_state_dict(sd)
    return m.float().eval(), {k: v.float() for k, v in sd.items()}


def _ocfg(d):
    return mo.OracleConfig(hidden_size=d.hidden_size, intermediate_size=d.intermediate_size,
                           num_hidden_layers=d.num_hidden_layers, num_attention_heads=d.num_attention_heads)


def _mk(n_ctx, n_gen, H, W, sp=1, seed=42):
    d = po.frame_block_inputs(n_ctx, n_gen, H, W, True, sp)
    lat = synth.synthetic_latents(n_ctx + n_gen, H, W, seed=seed)
    mk = dict(input_ids=d["input_ids"], input_img_latents=lat[:n_ctx], input_image_sizes=d["input_image_sizes"],
              attention_mask=d["attention_mask"], position_ids=d["position_ids"],
              denoise_image_sizes=d["denoise_image_sizes"], time_emb_inx=d["time_emb_inx"], img_cfg_scale=1.5,
              use_img_cfg=True, use_kv_cache=False, offload_model=False, vae=None)
    return mk, lat[n_ctx:]


def _maxerr(a, b):
    a, b = torch.cat([x.flatten() for x in a]), torch.cat([x.flatten() for x in b])
    return float((a - b).abs().max() / b.abs().max())


@pytest.mark.parametrize("geom", [(2, 2, 64, 64, 1), (3, 2, 64, 96, 1), (1, 1, 64, 64, 1), (3, 4, 48, 80, 8),
                                  (5, 2, 64, 96, 8)])
@pytest.mark.parametrize("pt", ["x1", "v"])
def test_engine_loop_reproduces_the_oracle_in_fp32(emu, geom, pt):
    """Fused scheduler loop (prefix cached once, unconditional row unpadded, codes instead of the
    mask) == the oracle's sampler over the reference path as written, 3 Euler steps with CFG; the
    left-padded / SP-padded layouts of the reference processor included."""
    from videogpt_b200 import LVMScheduler
    n_ctx, n_gen, H, W, sp = geom
    m, sd = _model()
    mk, z0 = _mk(n_ctx, n_gen, H, W, sp)
    with torch.no_grad():
        want = so.euler_sample([x.clone() for x in z0] * 2,
                               lambda z, t, **kw: mo.frame_block_forward_with_cfg(sd, _ocfg(synth.REDUCED), z, t, **kw),
                               mk, num_steps=3, prediction_type=pt)
    got = LVMScheduler(num_steps=3)([x.clone() for x in z0] * 2, m.frame_block_forward_with_cfg, mk,
                                    use_kv_cache=False, prediction_type=pt)
    assert _maxerr(got, want) < TOL
    assert emu.calls.count("attention") == synth.REDUCED.num_hidden_layers * 3 + (synth.REDUCED.num_hidden_layers - 1)


def test_callback_seam_reproduces_the_oracle_in_fp32(emu):
    """Seam S2: one call of frame_block_forward_with_cfg with per-latent timesteps (not uniform)."""
    m, sd = _model()
    mk, z0 = _mk(2, 3, 64, 64)
    z = [x.clone() for x in z0] * 2
    t = torch.tensor([0.1, 0.4, 0.7, 0.1, 0.4, 0.7])
    for pt in ("x1", "v"):
        with torch.no_grad():
            want = mo.frame_block_forward_with_cfg(sd, _ocfg(synth.REDUCED), z, t, prediction_type=pt, **mk)
        got, _ = m.frame_block_forward_with_cfg(z, t, past_key_values=None, prediction_type=pt, **mk)
        assert _maxerr(got, want) < TOL


@pytest.mark.parametrize("geom", [(2, 64, 64, 1), (3, 48, 80, 4), (1, 64, 64, 1)])
@pytest.mark.parametrize("pt", ["x1", "v"])
def test_single_frame_path_reproduces_the_oracle_in_fp32(emu, geom, pt):
    """``pipeline.__call__`` layout (LVM.forward_with_cfg): condition tokens cached as the prefix,
    [time | image] rows active; the 1-token unconditional row left-padded by the reference."""
    from videogpt_b200 import LVMScheduler
    n_ctx, H, W, sp = geom
    m, sd = _model()
    d = po.single_frame_inputs(n_ctx, H, W, True, sp)
    lat = synth.synthetic_latents(n_ctx + 1, H, W, seed=7)
    mk = dict(input_ids=d["input_ids"], input_img_latents=lat[:n_ctx], input_image_sizes=d["input_image_sizes"],
              attention_mask=d["attention_mask"], position_ids=d["position_ids"], img_cfg_scale=1.5,
              use_img_cfg=True, use_kv_cache=False, offload_model=False)
    z0 = torch.cat([lat[n_ctx]] * 2, 0)
    with torch.no_grad():
        want = so.euler_sample(z0.clone(),
                               lambda z, t, **kw: mo.single_frame_forward_with_cfg(sd, _ocfg(synth.REDUCED), z, t, **kw),
                               mk, num_steps=3, prediction_type=pt)
    got = LVMScheduler(num_steps=3)(z0.clone(), m.forward_with_cfg, mk, use_kv_cache=False, prediction_type=pt)
    assert _maxerr([got], [want]) < TOL


def test_next_clip_of_the_same_geometry_reuses_the_plan_and_redoes_the_prefill(emu):
    """Cache level 2 of LVM.prepare_frame_block: new context tensors, same layout -> same plan
    object, new context K/V; results equal the oracle for BOTH clips."""
    from videogpt_b200 import LVMScheduler
    m, sd = _model()
    plans = []
    for seed in (1, 2):
        mk, z0 = _mk(2, 2, 64, 64, seed=seed)
        with torch.no_grad():
            want = so.euler_sample([x.clone() for x in z0] * 2,
                                   lambda z, t, **kw: mo.frame_block_forward_with_cfg(sd, _ocfg(synth.REDUCED), z, t, **kw),
                                   mk, num_steps=2, prediction_type="x1")
        got = LVMScheduler(num_steps=2)([x.clone() for x in z0] * 2, m.frame_block_forward_with_cfg, mk,
                                        use_kv_cache=False, prediction_type="x1")
        assert _maxerr(got, want) < TOL
        plans.append(m._engine.plan)
    assert plans[0] is plans[1]


def _pipe(m):
    from videogpt_b200 import LVMPipeline, LVMProcessor
    return LVMPipeline(None, m, LVMProcessor(synth.SingleIdTagTokenizer()), device="cpu")


@pytest.mark.parametrize("guidance", [1.5, 1.0])
def test_batched_videos_reproduce_the_oracle_and_the_per_video_runs(emu, guidance):
    """BASELINE configs[3] (batch of videos x CFG branches in one pass): (a) the batched index
    dicts fed to the ORACLE -- the reference model as written takes any number of rows
    (LVM/model.py:436-453) -- give what the product computes from them; (b) video v of the batch
    equals video v run alone through the single-video API."""
    from videogpt_b200.pipeline import replicate_frame_block_inputs
    n_videos, n_ctx, n_gen, H, W, steps = 3, 2, 2, 64, 64, 2
    m, sd = _model()
    pipe = _pipe(m)
    lats = [synth.synthetic_latents(n_ctx + n_gen, H, W, seed=10 + v) for v in range(n_videos)]
    ctx = [l[:n_ctx] for l in lats]
    noise = [l[n_ctx:] for l in lats]
    kw = dict(num_inference_steps=steps, img_guidance_scale=guidance, prediction_type="x1", dtype=torch.float32)
    got = pipe.next_clip_latents_batch(ctx, n_gen, initial_noise=noise, **kw)
    assert len(got) == n_videos and all(len(g) == n_gen for g in got)
    # (b) per-video runs
    for v in range(n_videos):
        alone = pipe.next_clip_latents(ctx[v], n_gen, initial_noise=noise[v], **kw)
        assert _maxerr(got[v], alone) < TOL
    # (a) the oracle on the batched rows, dense mask
    use_cfg = guidance != 1.0
    d = replicate_frame_block_inputs(po.frame_block_inputs(n_ctx, n_gen, H, W, use_cfg, 1), n_videos)
    mk = dict(input_ids=d["input_ids"], input_img_latents=[x for c in ctx for x in c],
              input_image_sizes=d["input_image_sizes"], attention_mask=d["attention_mask"],
              position_ids=d["position_ids"], denoise_image_sizes=d["denoise_image_sizes"],
              time_emb_inx=d["time_emb_inx"], img_cfg_scale=guidance, use_img_cfg=use_cfg)
    z0 = [x.clone() for nv in noise for x in nv] * (2 if use_cfg else 1)
    with torch.no_grad():
        want = so.euler_sample(z0, lambda z, t, **kw_: mo.frame_block_forward_with_cfg(sd, _ocfg(synth.REDUCED), z, t, **kw_),
                               mk, num_steps=steps, prediction_type="x1")
    assert _maxerr([x for g in got for x in g], want[:n_videos * n_gen]) < TOL


def test_pipeline_call_one_frame_at_a_time_reproduces_the_oracle(emu):
    """``LVMPipeline.__call__`` (reference pipeline.py:136-343): two input images, two generated
    frames; the latent handed to the VAE decoder for every generated frame must be what the
    oracle's sampler gives for the same context latents and the same seeded noise."""
    m, sd = _model()
    pipe = _pipe(m)
    pipe.vae = vae = _FakeVAE()
    imgs = [_pil(1), _pil(2)]
    out = pipe(input_images=imgs, height=64, width=64, gen_num=2, num_inference_steps=2, img_guidance_scale=1.5,
               use_input_image_size_as_output=True, dtype=torch.float32, seed=5, prediction_type="v",
               clean_image_noise_level=0.0)
    assert len(out) == 4 and all(im.size == (64, 64) for im in out)
    # decode order: 2 context reconstructions, generated frame 1, generated frame 2
    assert len(vae.decoded) == 4
    ctx = [pipe.vae_encode(pipe.processor.process_image(im).unsqueeze(0), torch.float32) for im in imgs]
    for k in range(2):
        if k == 1:     # the second frame is conditioned on the first one, re-encoded from its PIL image
            ctx.append(pipe.vae_encode(pipe.processor.process_image(out[2]).unsqueeze(0), torch.float32))
        d = po.single_frame_inputs(len(ctx), 64, 64, True, 1)
        mk = dict(input_ids=d["input_ids"], input_img_latents=ctx, input_image_sizes=d["input_image_sizes"],
                  attention_mask=d["attention_mask"], position_ids=d["position_ids"], img_cfg_scale=1.5,
                  use_img_cfg=True)
        noise = torch.randn(1, 4, 8, 8, generator=torch.Generator().manual_seed(5))
        with torch.no_grad():
            want = so.euler_sample(torch.cat([noise] * 2, 0),
                                   lambda z, t, **kw: mo.single_frame_forward_with_cfg(sd, _ocfg(synth.REDUCED), z, t, **kw),
                                   mk, num_steps=2, prediction_type="v")[:1]
        got = vae.decoded[2 + k] * vae.config.scaling_factor
        assert _maxerr([got], [want]) < TOL


def test_frame_block_autoregressive_entry_point_reproduces_the_oracle(emu):
    """``LVMPipeline.prompt_condition_frame_block_autoregressive_inference`` (reference pipeline.py:346-595, the entry
    point of the shipped inference script): two rounds of two generated frames from three context images through a
    fake VAE.  Every latent handed to the VAE decoder must be what the oracle's frame-block sampler gives for the same
    context latents (re-encoded from the PIL frames of the previous round, like the reference) and the same seeded noise;
    the second round also exercises the ``max_frame_window`` cut (pipeline.py:421-422)."""
    m, sd = _model()
    pipe = _pipe(m)
    pipe.vae = vae = _FakeVAE()
    imgs = [_pil(11), _pil(12), _pil(13)]
    out = pipe.prompt_condition_frame_block_autoregressive_inference(
        input_images=imgs, height=64, width=64, gen_nums=[2, 2], num_inference_steps=2, img_guidance_scale=1.5,
        use_input_image_size_as_output=True, dtype=torch.float32, seed=9, prediction_type="x1",
        clean_image_noise_level=0.0, max_frame_window=6)
    assert len(out) == 3 + 2 + 2 and all(im.size == (64, 64) for im in out)
    assert len(vae.decoded) == 7                     # 3 context reconstructions, then 2 + 2 generated frames
    frames = list(imgs)
    for k in range(2):
        if k == 1:
            frames = out[:5][-4:]                    # window 6 - 2 generated = the last 4 frames so far
        ctx = [pipe.vae_encode(pipe.processor.process_image(im).unsqueeze(0), torch.float32) for im in frames]
        d = po.frame_block_inputs(len(ctx), 2, 64, 64, True, 1)
        mk = dict(input_ids=d["input_ids"], input_img_latents=ctx, input_image_sizes=d["input_image_sizes"],
                  attention_mask=d["attention_mask"], position_ids=d["position_ids"],
                  denoise_image_sizes=d["denoise_image_sizes"], time_emb_inx=d["time_emb_inx"], img_cfg_scale=1.5,
                  use_img_cfg=True)
        g = torch.Generator().manual_seed(9)
        noise = [torch.randn(1, 4, 8, 8, generator=g) for _ in range(2)]
        with torch.no_grad():
            want = so.euler_sample(noise * 2, lambda z, t, **kw: mo.frame_block_forward_with_cfg(sd, _ocfg(synth.REDUCED), z, t, **kw),
                                   mk, num_steps=2, prediction_type="x1")[:2]
        got = [vae.decoded[3 + 2 * k + i] * vae.config.scaling_factor for i in range(2)]
        assert _maxerr(got, want) < TOL, k


def test_pipeline_call_without_input_images_generates_unconditionally(emu):
    """No input images: guidance is switched off (pipeline.py:217-218), the sequence is
    ``[<|diffusion|>, time, image tokens]``; the second frame is conditioned on the first."""
    m, sd = _model()
    pipe = _pipe(m)
    pipe.vae = vae = _FakeVAE()
    out = pipe(input_images=None, height=64, width=64, gen_num=2, num_inference_steps=2, dtype=torch.float32,
               seed=3, prediction_type="x1", clean_image_noise_level=0.0)
    assert len(out) == 2 and len(vae.decoded) == 2      # nothing to reconstruct: the two generated frames
    d = po.single_frame_inputs(0, 64, 64, False, 1)
    mk = dict(input_ids=d["input_ids"], input_img_latents=[], input_image_sizes=d["input_image_sizes"],
              attention_mask=d["attention_mask"], position_ids=d["position_ids"], img_cfg_scale=1.6, use_img_cfg=False)
    noise = torch.randn(1, 4, 8, 8, generator=torch.Generator().manual_seed(3))
    with torch.no_grad():
        want = so.euler_sample(noise, lambda z, t, **kw: mo.single_frame_forward_with_cfg(sd, _ocfg(synth.REDUCED), z, t, **kw),
                               mk, num_steps=2, prediction_type="x1")
    assert _maxerr([vae.decoded[0] * vae.config.scaling_factor], [want]) < TOL


# ---- latent-space rollout with a persistent paged K/V cache ---------------------
def _history_mask(d, n_hist, gen, bl, ws_of_frame, ws_now):
    """The reference mask over the FULL history (cond row), minus what the rollout's window hides:
    the clip sees context frames >= ws_now, context frame f keeps the view of the round that cached
    it (frames >= ws_of_frame[f]).  Row 1 (unconditional) is untouched."""
    mask = d["attention_mask"].clone()
    frame = torch.arange(mask.shape[-1]) // bl
    for q in range(mask.shape[-1]):
        fq = int(frame[q])
        lo = ws_now if fq >= n_hist else ws_of_frame[fq]
        mask[0, q, frame < lo] = 0
    return mask


@pytest.mark.parametrize("pt", ["x1", "v"])
def test_rollout_with_persistent_cache_matches_the_oracle_over_the_full_history(emu, pt):
    """4 rounds, window 6 frames, 2 frames per clip, 128x128 frames (66-token blocks, so evicted
    frames release whole 128-token pages): every round's clip equals the oracle's sampler over
    the full history with absolute positions and the history mask; every context frame goes
    through the transformer once; from round 2 on the plan is refreshed in place."""
    from videogpt_b200 import LVMProcessor
    from videogpt_b200.rollout import LatentRollout, window_start
    n0, gen, window, H, W, steps, rounds = 2, 2, 6, 128, 128, 2, 4
    bl = (H // 16) * (W // 16) + 2
    m, sd = _model()
    lat = synth.synthetic_latents(n0 + gen * rounds, H, W, seed=3)
    history = [x.clone() for x in lat[:n0]]
    noises = [lat[n0 + gen * r:n0 + gen * (r + 1)] for r in range(rounds)]
    ro = LatentRollout(m, LVMProcessor(synth.SingleIdTagTokenizer()), gen, max_frame_window=window,
                       num_inference_steps=steps, img_guidance_scale=1.5, prediction_type=pt).start(history)
    ws_of_frame, plans, freed = {}, [], False
    for r in range(rounds):
        n_hist = len(history)
        ws = window_start(n_hist, gen, window)
        for f in range(n_hist):
            ws_of_frame.setdefault(f, ws)
        got = ro.next_clip(initial_noise=noises[r])
        plans.append(ro.engine.plan)
        freed = freed or min(ro._phys) > 0
        d = po.frame_block_inputs(n_hist, gen, H, W, True, 1)
        mk = dict(input_ids=d["input_ids"], input_img_latents=history, input_image_sizes=d["input_image_sizes"],
                  attention_mask=_history_mask(d, n_hist, gen, bl, ws_of_frame, ws), position_ids=d["position_ids"],
                  denoise_image_sizes=d["denoise_image_sizes"], time_emb_inx=d["time_emb_inx"], img_cfg_scale=1.5,
                  use_img_cfg=True)
        with torch.no_grad():
            want = so.euler_sample([x.clone() for x in noises[r]] * 2,
                                   lambda z, t, **kw: mo.frame_block_forward_with_cfg(sd, _ocfg(synth.REDUCED), z, t, **kw),
                                   mk, num_steps=steps, prediction_type=pt)[:gen]
        assert _maxerr(got, want) < TOL, f"round {r}"
        history += [x.clone() for x in got]
    assert ro.prefilled_frames == n0 + gen * (rounds - 1)          # every context frame exactly once
    assert plans[1] is plans[2] is plans[3]                          # refreshed in place (graph stays valid)
    assert freed                                                     # pages behind the window were released


def test_rollout_equals_the_reference_round_while_the_window_holds_the_history(emu):
    """No eviction: a rollout round IS the reference's round (pipeline.next_clip_latents on all
    frames so far) -- same mask, same positions -- although only the new frames are prefilled."""
    from videogpt_b200 import LVMProcessor
    from videogpt_b200.rollout import LatentRollout
    n0, gen, H, W, steps = 1, 2, 64, 64, 2
    m, sd = _model()
    pipe = _pipe(m)
    lat = synth.synthetic_latents(n0 + 3 * gen, H, W, seed=9)
    history = [x.clone() for x in lat[:n0]]
    ro = LatentRollout(m, pipe.processor, gen, max_frame_window=16, num_inference_steps=steps,
                       img_guidance_scale=1.5, prediction_type="x1").start(history)
    outs = []
    for r in range(3):
        outs.append(ro.next_clip(initial_noise=lat[n0 + gen * r:n0 + gen * (r + 1)]))
    for r in range(3):        # afterwards, so that the rollout's cache is not disturbed by these calls
        want = pipe.next_clip_latents(history, gen, num_inference_steps=steps, img_guidance_scale=1.5,
                                      prediction_type="x1", dtype=torch.float32,
                                      initial_noise=lat[n0 + gen * r:n0 + gen * (r + 1)])
        assert _maxerr(outs[r], want) < TOL
        history += [x.clone() for x in outs[r]]


def test_pipeline_rollout_latents_persistent_and_recomputed_agree_without_eviction(emu):
    m, _ = _model()
    pipe = _pipe(m)
    ctx = synth.synthetic_latents(2, 64, 64, seed=4)
    kw = dict(num_inference_steps=2, img_guidance_scale=1.5, seed=11, prediction_type="v", dtype=torch.float32)
    a = pipe.rollout_latents(ctx, [2, 2, 2], persistent_cache=True, **kw)
    b = pipe.rollout_latents(ctx, [2, 2, 2], persistent_cache=False, **kw)
    assert len(a) == len(b) == 6 and _maxerr(a, b) < TOL
    # with eviction the two differ by design (windowed attention over cached K/V vs recompute)
    c = pipe.rollout_latents(ctx, [2, 2, 2], persistent_cache=True, max_frame_window=4, **kw)
    assert _maxerr(c[:2], a[:2]) < TOL and all(torch.isfinite(x).all() for x in c)


def test_rollout_refuses_to_continue_after_the_engine_was_used_for_another_clip(emu):
    from videogpt_b200.rollout import LatentRollout
    m, _ = _model()
    pipe = _pipe(m)
    ctx = synth.synthetic_latents(2, 64, 64, seed=4)
    ro = LatentRollout(m, pipe.processor, 2, num_inference_steps=1, prediction_type="x1").start(ctx)
    ro.next_clip(seed=1)
    pipe.next_clip_latents(ctx, 2, num_inference_steps=1, dtype=torch.float32, seed=1)     # replaces the engine's plan
    with pytest.raises(RuntimeError, match="start\\(\\) again"):
        ro.next_clip(seed=1)


# ---- sequence parallelism: row-sharded plans on virtual ranks --------------------------------------
@pytest.mark.parametrize("world,geom,partition", [(2, (3, 2, 64, 96), "rows"), (3, (2, 2, 64, 64), "rows"), (4, (8, 2, 64, 64), "rows"),
                                                  (8, (4, 4, 128, 128), "rows"), (8, (1, 1, 32, 32), "rows"),
                                                  (2, (2, 2, 256, 256), "rows"), (3, (3, 3, 256, 256), "rows"),   # two ranges per rank
                                                  (2, (3, 2, 64, 96), "sequences"), (2, (1, 1, 32, 32), "sequences")])
def test_row_sharded_virtual_ranks_reproduce_the_unsharded_engine_and_the_oracle(emu, world, geom, partition):
    """CPU twin of tests/test_sequence_parallel.py::test_virtual_ranks_match_unsharded_engine_bit_exact:
    every virtual rank runs its chunk of the rows of every sequence, the K/V-append and final-layer
    emulations store into every rank's buffers through the raw pointers the engine hands out, and
    all ranks end up with the complete K/V pool and the complete prediction -- equal to the
    unsharded engine's (fp32 rounding) and, over 3 Euler steps, to the oracle's sampler.
    ``partition="sequences"``: one CFG branch per rank -- K/V stay on the rank that owns the sequence,
    only the prediction is stored into the peer."""
    from videogpt_b200 import engine as eng, peer
    n_ctx, n_gen, H, W = geom
    dims = synth.REDUCED
    sd = synth.init_state_dict(dims, seed=0, dtype=torch.float32)
    dev = torch.device("cpu")
    w = eng.EngineWeights(sd, dims.num_hidden_layers, dev)
    d = po.frame_block_inputs(n_ctx, n_gen, H, W, True, 1)
    specs, n_lat, n_ctx_lat = eng.frame_block_specs(d["input_ids"], d["position_ids"], d["input_image_sizes"],
                                                    d["denoise_image_sizes"], d["time_emb_inx"])
    lat = synth.synthetic_latents(n_ctx + n_gen, H, W, seed=42)
    ctx, z0 = torch.cat(lat[:n_ctx], 0), torch.cat(lat[n_ctx:] * 2, 0)

    def engine(peers=None):
        return eng.NextClipEngine(w, dims.hidden_size, dims.intermediate_size, dims.num_hidden_layers,
                                  dims.num_attention_heads, dims.rms_norm_eps, dims.rope_theta, dev,
                                  use_cuda_graph=False, peers=peers)

    ref = engine()
    ref.set_plan(eng.build_plan(specs, n_lat, n_ctx_lat, H // 8, W // 8, dev))
    ref.prefill(ctx)
    ranks = [engine(m) for m in peer.LocalPeerGroup.create(world, dev)]
    for r, e in enumerate(ranks):
        e.set_plan(eng.build_plan(specs, n_lat, n_ctx_lat, H // 8, W // 8, dev, shard=(r, world), partition=partition))
        emu.register_peer_buffers(e._kv_shared)
        emu.register_peer_buffers(e._pred_shared)
    eng.run_lockstep([e.prefill_steps(ctx) for e in ranks])
    for r, e in enumerate(ranks):
        if partition == "rows":
            assert (e.kv - ref.kv).abs().max() < 1e-4
        else:           # only its own sequence's pages; nobody wrote the other sequence's
            mine = ref.plan.page_table[r, :(specs[r].n_prefix + specs[r].n_active + 127) // 128].long()
            other = ref.plan.page_table[1 - r, :(specs[1 - r].n_prefix + specs[1 - r].n_active + 127) // 128].long()
            assert (e.kv[:, :, mine] - ref.kv[:, :, mine]).abs().max() < 1e-4 and e.kv[:, :, other].abs().max() == 0
    sigma = torch.linspace(0, 1, 4)
    for e in [ref] + ranks:
        e.z.copy_(z0)
    for i in range(3):
        for e in [ref] + ranks:
            e.t.fill_(float(sigma[i]))
        ref.predict()
        eng.run_lockstep([e.predict_steps() for e in ranks])
        for e in ranks:
            assert (e.pred - ref.pred).abs().max() < 1e-4 * float(ref.pred.abs().max()), f"step {i}"
        for e in [ref] + ranks:
            emu.cfg_euler(e.z, e.pred, True, True, float(1 - sigma[i]), float(sigma[i + 1] - sigma[i]), 1.5)
    mk, z_list = _mk(n_ctx, n_gen, H, W)
    with torch.no_grad():
        want = so.euler_sample([x.clone() for x in z_list] * 2,
                               lambda z, t, **kw: mo.frame_block_forward_with_cfg(sd, _ocfg(dims), z, t, **kw),
                               mk, num_steps=3, prediction_type="x1")
    for e in ranks:
        assert _maxerr([e.z], [torch.cat(want, 0)]) < 5 * TOL


def _ranks_as_threads(emu, world, partition, body):
    """Run ``body(rank, model, counts)`` on ``world`` ranks as threads: every rank gets its own model whose engine is
    a member of one peer group (barrier = threading.Barrier, peer buffers = every rank's buffers registered with the
    kernel emulation), under ``hccl_info.partition = partition``.  ``counts[rank]`` counts the data-path barriers."""
    import threading
    from videogpt_b200 import engine as eng, parallel_states as ps, peer
    bar, lock = threading.Barrier(world), threading.Lock()
    counts = [0] * world

    class ThreadPeers(peer.LocalPeerGroup):
        lockstep = False

        def alloc(self, nbytes):
            with lock:
                buf = super().alloc(nbytes)
                emu.register_peer_buffers(buf)
            bar.wait()
            return buf

        def barrier(self):
            counts[self.rank] += 1
            bar.wait()

        def host_barrier(self):
            bar.wait()

    registry = []
    members = [ThreadPeers(r, world, torch.device("cpu"), registry) for r in range(world)]
    models = [_model()[0] for _ in range(world)]
    results, errors = [None] * world, []

    def rank_main(r):
        try:
            m, d = models[r], models[r].dims()
            m._engine = eng.NextClipEngine(eng.EngineWeights(m.state_dict(), d.num_hidden_layers, "cpu"), d.hidden_size,
                                           d.intermediate_size, d.num_hidden_layers, d.num_attention_heads,
                                           d.rms_norm_eps, d.rope_theta, "cpu", use_cuda_graph=False, peers=members[r])
            m._engine_key = tuple(p._version for p in m.parameters())
            m.engine = lambda: m._engine
            results[r] = body(r, m, counts)
        except BaseException as exc:          # a dead rank must not leave the others waiting forever
            errors.append(exc)
            bar.abort()

    ps.hccl_info.partition = partition
    try:
        threads = [threading.Thread(target=rank_main, args=(r,)) for r in range(world)]
        for t in threads:
            t.start()
        for t in threads:
            t.join(300)
    finally:
        ps.hccl_info.partition = "rows"
    assert not errors, errors
    return results


def test_cfg_branch_pair_through_model_and_scheduler_reproduces_the_oracle(emu):
    """The user-facing flow of a CFG-branch pair: ``initialize_sequence_parallel_state(2, partition="sequences")``
    (here: its state set by hand, two ranks as threads), then ``LVMScheduler`` on ``frame_block_forward_with_cfg``
    as on one GPU.  Barriers per clip: none in the prefill, two per Euler step; both ranks end on the oracle's
    latents, x1 and v, and a second clip with fresh context tensors reuses the plan."""
    from videogpt_b200 import LVMScheduler
    world, n_ctx, n_gen, H, W, steps = 2, 3, 2, 64, 96, 3
    sd = _model()[1]

    def body(r, m, counts):
        got = {}
        for pt in ("x1", "v"):
            for clip in range(2):
                mk, z_list = _mk(n_ctx, n_gen, H, W)
                before = counts[r]
                out = LVMScheduler(steps)([x.clone() for x in z_list] * 2, m.frame_block_forward_with_cfg, mk,
                                          prediction_type=pt)
                got[(pt, clip)] = (out, counts[r] - before)
        return got, (m._engine.plan.partition, m._engine.plan.prefix.rows, m._engine.plan.step.rows)

    res = _ranks_as_threads(emu, world, "sequences", body)
    block = H * W // 256 + 2
    assert [x[1] for x in res] == [("sequences", n_ctx * block, n_gen * block), ("sequences", 0, n_gen * block)]
    for pt in ("x1", "v"):
        mk, z_list = _mk(n_ctx, n_gen, H, W)
        with torch.no_grad():
            want = so.euler_sample([x.clone() for x in z_list] * 2,
                                   lambda z, t, **kw: mo.frame_block_forward_with_cfg(sd, _ocfg(synth.REDUCED), z, t, **kw),
                                   mk, num_steps=steps, prediction_type=pt)
        for r in range(world):
            for clip in range(2):
                out, n_barriers = res[r][0][(pt, clip)]
                assert n_barriers == 2 * steps, (r, pt, clip, n_barriers)
                assert _maxerr(out, want) < 5 * TOL, (r, pt, clip)


@pytest.mark.parametrize("world,partition,geom", [(2, "sequences", (2, 64, 64, 1)), (2, "rows", (3, 48, 80, 4)),
                                                  (3, "rows", (2, 64, 64, 1))])
def test_single_frame_path_in_a_peer_group_reproduces_the_oracle(emu, world, partition, geom):
    """``pipeline.__call__``'s layout (``LVM.forward_with_cfg``: condition tokens as the cached prefix, [time | image]
    rows active, a one-token unconditional row) with the rows of every sequence dealt to the ranks, and with one CFG
    branch per rank: every rank ends on the oracle's latents."""
    from videogpt_b200 import LVMScheduler
    n_ctx, H, W, sp = geom
    sd = _model()[1]
    d = po.single_frame_inputs(n_ctx, H, W, True, sp)
    lat = synth.synthetic_latents(n_ctx + 1, H, W, seed=7)
    z0 = torch.cat([lat[n_ctx]] * 2, 0)

    def mk():
        return dict(input_ids=d["input_ids"], input_img_latents=[x.clone() for x in lat[:n_ctx]],
                    input_image_sizes=d["input_image_sizes"], attention_mask=d["attention_mask"],
                    position_ids=d["position_ids"], img_cfg_scale=1.5, use_img_cfg=True, use_kv_cache=False,
                    offload_model=False)

    with torch.no_grad():
        want = so.euler_sample(z0.clone(),
                               lambda z, t, **kw: mo.single_frame_forward_with_cfg(sd, _ocfg(synth.REDUCED), z, t, **kw),
                               mk(), num_steps=3, prediction_type="x1")

    def body(r, m, counts):
        got = LVMScheduler(num_steps=3)(z0.clone(), m.forward_with_cfg, mk(), use_kv_cache=False, prediction_type="x1")
        return got, m._engine.plan.partition, m._engine.plan.step.rows

    res = _ranks_as_threads(emu, world, partition, body)
    n_tok = H * W // 256
    assert sum(x[2] for x in res) == 2 * (n_tok + 1) and all(x[1] == partition for x in res)
    if partition == "sequences":
        assert [x[2] for x in res] == [n_tok + 1, n_tok + 1]
    for r in range(world):
        assert _maxerr([res[r][0]], [want]) < 5 * TOL, r


def test_rollout_under_sequence_parallelism_matches_the_single_rank_rollout(emu):
    """BASELINE configs[2] in full: long-context rollout with the persistent K/V cache AND the rows
    of every round dealt to the ranks of a sequence-parallel group.  Three virtual ranks run as
    threads (barrier = threading.Barrier; K/V and predictions stored into every rank's buffers by
    the emulated peer kernels); with eviction, 4 rounds; every rank must reproduce the single-rank
    rollout, which the oracle test above pins."""
    import threading
    from videogpt_b200 import LVMProcessor, peer
    from videogpt_b200.rollout import LatentRollout
    world, n0, gen, window, H, W, steps, rounds = 3, 2, 2, 6, 128, 128, 2, 4
    lat = synth.synthetic_latents(n0 + gen * rounds, H, W, seed=3)
    noises = [lat[n0 + gen * r:n0 + gen * (r + 1)] for r in range(rounds)]
    proc = LVMProcessor(synth.SingleIdTagTokenizer())
    kw = dict(max_frame_window=window, num_inference_steps=steps, img_guidance_scale=1.5, prediction_type="x1")

    m0, _ = _model()
    single = LatentRollout(m0, proc, gen, **kw).start(lat[:n0])
    want = [single.next_clip(initial_noise=noises[r]) for r in range(rounds)]

    bar, lock = threading.Barrier(world), threading.Lock()

    class ThreadPeers(peer.LocalPeerGroup):
        lockstep = False

        def alloc(self, nbytes):
            with lock:
                buf = super().alloc(nbytes)
                emu.register_peer_buffers(buf)
            bar.wait()
            return buf

        def barrier(self):
            bar.wait()

        host_barrier = barrier

    registry = []
    members = [ThreadPeers(r, world, torch.device("cpu"), registry) for r in range(world)]
    models = [_model()[0] for _ in range(world)]
    got, errors = [None] * world, []

    def rank_main(r):
        try:
            from videogpt_b200 import engine as eng
            m, d = models[r], models[r].dims()
            m._engine = eng.NextClipEngine(eng.EngineWeights(m.state_dict(), d.num_hidden_layers, "cpu"), d.hidden_size,
                                           d.intermediate_size, d.num_hidden_layers, d.num_attention_heads,
                                           d.rms_norm_eps, d.rope_theta, "cpu", use_cuda_graph=False, peers=members[r])
            m._engine_key = tuple(p._version for p in m.parameters())
            ro = LatentRollout(m, proc, gen, **kw).start(lat[:n0])
            got[r] = [ro.next_clip(initial_noise=noises[k]) for k in range(rounds)]
            assert ro.engine.plan.shard == (r, world)
        except BaseException as exc:          # a dead rank must not leave the others waiting forever
            errors.append(exc)
            bar.abort()

    threads = [threading.Thread(target=rank_main, args=(r,)) for r in range(world)]
    for t in threads:
        t.start()
    for t in threads:
        t.join(300)
    assert not errors, errors
    for r in range(world):
        for k in range(rounds):
            assert _maxerr(got[r][k], want[k]) < 5 * TOL, f"rank {r} round {k}"


def test_rollout_window_smaller_than_two_clips_skips_frames_that_are_never_visible(emu):
    """max_frame_window - gen_num < gen_num: part of a generated clip is already outside the window
    when it becomes context -- those frames are never prefilled (and never visible), the rest is."""
    from videogpt_b200 import LVMProcessor
    from videogpt_b200.rollout import LatentRollout, window_start
    n0, gen, window, H, W, steps, rounds = 1, 3, 4, 64, 64, 2, 3
    bl = (H // 16) * (W // 16) + 2
    m, sd = _model()
    lat = synth.synthetic_latents(n0 + gen * rounds, H, W, seed=8)
    history = [x.clone() for x in lat[:n0]]
    ro = LatentRollout(m, LVMProcessor(synth.SingleIdTagTokenizer()), gen, max_frame_window=window,
                       num_inference_steps=steps, img_guidance_scale=1.5, prediction_type="x1").start(history)
    ws_of_frame = {}
    for r in range(rounds):
        n_hist = len(history)
        ws = window_start(n_hist, gen, window)
        for f in range(n_hist):
            ws_of_frame.setdefault(f, min(ws, f))        # a frame outside the window keeps (only) itself in the oracle
        noise = lat[n0 + gen * r:n0 + gen * (r + 1)]
        got = ro.next_clip(initial_noise=noise)
        d = po.frame_block_inputs(n_hist, gen, H, W, True, 1)
        mk = dict(input_ids=d["input_ids"], input_img_latents=history, input_image_sizes=d["input_image_sizes"],
                  attention_mask=_history_mask(d, n_hist, gen, bl, ws_of_frame, ws), position_ids=d["position_ids"],
                  denoise_image_sizes=d["denoise_image_sizes"], time_emb_inx=d["time_emb_inx"], img_cfg_scale=1.5,
                  use_img_cfg=True)
        with torch.no_grad():
            want = so.euler_sample([x.clone() for x in noise] * 2,
                                   lambda z, t, **kw: mo.frame_block_forward_with_cfg(sd, _ocfg(synth.REDUCED), z, t, **kw),
                                   mk, num_steps=steps, prediction_type="x1")[:gen]
        assert _maxerr(got, want) < TOL, f"round {r}"
        history += [x.clone() for x in got]
    assert ro.prefilled_frames == n0 + 1 + 1            # per later round only the one frame inside the window


def test_model_helper_methods_match_the_oracle(emu):
    """The reference model's public helpers (LVM/model.py:255-327) on the drop-in: ``cropped_pos_embed``
    and ``unpatchify`` are index arithmetic; ``patch_multiple_resolutions`` goes through the assembly
    kernel (emulated here) for both embedders, tensors and lists."""
    m, sd = _model()
    cfg = _ocfg(synth.REDUCED)
    lat = torch.cat(synth.synthetic_latents(3, 64, 96, seed=2), 0)                        # [3,4,8,12]
    pe = m.cropped_pos_embed(8, 12)
    assert torch.equal(pe, mo.cropped_pos_embed(sd["pos_embed"], cfg.pos_embed_max_size, 8, 12, 2))
    for is_ctx, name in ((False, "x_embedder"), (True, "input_x_embedder")):
        got, n_tok, shapes = m.patch_multiple_resolutions(lat, is_input_images=is_ctx)
        want = torch.cat([mo.patch_embed(lat[i:i + 1], sd[f"{name}.proj.weight"], sd[f"{name}.proj.bias"],
                                         sd["pos_embed"], cfg) for i in range(3)], 0)
        assert n_tok == 24 and shapes == [8, 12] and float((got - want).abs().max()) < 1e-5
        got_l, n_l, shapes_l = m.patch_multiple_resolutions([lat[:1], lat[1:3]], is_input_images=is_ctx)
        assert n_l == [24, 24] and shapes_l == [[8, 12], [8, 12]] and float((torch.cat(got_l, 0) - want).abs().max()) < 1e-5
    y = torch.randn(2, 24, 16)
    assert torch.equal(m.unpatchify(y, 8, 12), mo.unpatchify(y, 8, 12, cfg))
    with pytest.raises(ValueError, match="pos_embed_max_size"):
        m.cropped_pos_embed(2 * 193, 8)


def test_submodule_forwards_match_the_oracle(emu):
    """TimestepEmbedder / PatchEmbedMR / FinalLayer of the drop-in are callable like the reference's
    modules (LVM/model.py:26-83, 138-154) and run on the hot path's kernels (emulated here)."""
    m, sd = _model()
    cfg = _ocfg(synth.REDUCED)
    t = torch.tensor([0.0, 0.25, 0.9])
    for name, mod in (("time_token", m.time_token), ("t_embedder", m.t_embedder)):
        assert float((mod(t) - mo.timestep_embedder(sd, name, t, torch.float32)).abs().max()) < 1e-5
    lat = torch.cat(synth.synthetic_latents(2, 64, 96, seed=5), 0)
    want = torch.nn.functional.conv2d(lat, sd["x_embedder.proj.weight"], sd["x_embedder.proj.bias"], stride=2)
    assert float((m.x_embedder(lat) - want.flatten(2).transpose(1, 2)).abs().max()) < 1e-5
    x, c = torch.randn(2, 24, synth.REDUCED.hidden_size), torch.randn(2, synth.REDUCED.hidden_size)
    assert float((m.final_layer(x, c) - mo.final_layer(sd, x, c)).abs().max()) < 1e-4
