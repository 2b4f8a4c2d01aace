"""Host plumbing dry run on CPU: scheduler -> model -> engine -> ops call sequence with the CUDA
entry points replaced by shape-checking stubs.  No arithmetic is checked here (that is the job of
the -m gpu parity tests); this only proves the host code builds plans, caches them, and calls the
C-ABI wrappers with consistent shapes and in the documented order."""
import pytest
import torch

from oracle import processor_oracle as po
from videogpt_b200 import synth

BF = torch.bfloat16


class StubOps:
    """Stands in for videogpt_b200.ops; records calls, validates shapes, writes zeros."""
    EPI_STORE, EPI_RESIDUAL, EPI_SWIGLU = 0, 1, 2
    ROW_TOKEN, ROW_TIME, ROW_NOISY_PATCH, ROW_CONTEXT_PATCH = 0, 1, 2, 3
    PAGE_TOKENS, ATTN_KV_TILE = 128, 64

    def __init__(self):
        self.calls = []

    def _log(self, name):
        self.calls.append(name)

    def pack_gate_up(self, w):
        self._log("pack_gate_up"); return w.clone()

    def rope_table(self, inv_freq, max_pos, head_dim):
        self._log("rope_table"); return torch.zeros(max_pos, head_dim, dtype=BF)

    def timestep_sinusoid(self, t, freqs, out=None):
        self._log("timestep_sinusoid"); assert out.shape == (t.numel(), 256); return out

    def linear_small(self, x, w, bias, pre_silu=False, post_silu=False, out=None):
        self._log("linear_small"); assert x.shape[1] == w.shape[1] and out.shape == (x.shape[0], w.shape[0]); return out

    def embed_assemble(self, hidden, kind, a, b, *rest):
        self._log("embed_assemble"); assert kind.numel() == a.numel() == b.numel() == hidden.shape[0]; return hidden

    def rmsnorm(self, x, w, eps, out=None):
        self._log("rmsnorm"); assert out.shape == x.shape and w.numel() == x.shape[1]; return out

    def gemm(self, a, w, out=None, residual=None, epilogue=0, block_n=0, tail_mode=-1):
        self._log("gemm")
        n_out = w.shape[0] // 2 if epilogue == 2 else w.shape[0]
        assert a.shape[1] == w.shape[1] and out.shape == (a.shape[0], n_out)
        assert (residual is not None) == (epilogue == 1)
        return out

    def rope_kv_append(self, qkv, row_pos, row_slot, table, k_pool, v_pool, heads, head_dim):
        self._log("rope_kv_append")
        assert row_pos.numel() == row_slot.numel() == qkv.shape[0]
        assert int(row_pos.max()) < table.shape[0] and int(row_slot.max()) < k_pool.shape[0] * 128

    def attention(self, q, out, k_pool, v_pool, page_table, seqs, max_q_rows, q_code, k_code, mm, heads, head_dim, scale):
        self._log("attention")
        assert q.shape == out.shape and q_code.numel() == q.shape[0]
        assert int(seqs[:, 1].sum()) == q.shape[0] and int(seqs[:, 1].max()) == max_q_rows
        return out

    def final_layer(self, hidden, lat_row0, mod, w, bias, pred, norm_weight=None, rms_eps=0.0, euler=None):
        self._log("final_layer"); assert mod.shape[0] == pred.shape[0] and norm_weight.numel() == hidden.shape[1]
        if euler is not None:
            self._log("euler_fused"); assert euler[0].shape == pred.shape and euler[1].numel() == 3
        pred.zero_(); return pred

    def cfg_euler(self, z, pred, use_cfg, x1_mode, *a, **k):
        self._log("cfg_euler"); assert z.shape == pred.shape; return z

    def cfg_combine(self, pred, guidance):
        self._log("cfg_combine"); return pred

    def mask_from_codes(self, qc, kc):
        self._log("mask_from_codes"); return (qc[:, None] >= kc[None, :]).to(torch.uint8)


@pytest.fixture
def dry(monkeypatch):
    from videogpt_b200 import engine, model, scheduler
    stub = StubOps()
    for mod in (engine, model, scheduler):
        monkeypatch.setattr(mod, "ops", stub)
    monkeypatch.setattr(torch.cuda, "is_available", lambda: True)
    # the engine refuses CPU parameters; lift only that guard for the dry run
    orig_engine = model.LVM.engine

    def engine_cpu(self):
        dev = torch.device("cpu")
        if self._engine is None:
            d = self.dims()
            w = engine.EngineWeights(self.state_dict(), d.num_hidden_layers, dev)
            self._engine = engine.NextClipEngine(w, d.hidden_size, d.intermediate_size, d.num_hidden_layers,
                                                 d.num_attention_heads, d.rms_norm_eps, d.rope_theta, dev,
                                                 self.pos_embed_max_size, self.patch_size, use_cuda_graph=False)
        return self._engine
    monkeypatch.setattr(model.LVM, "engine", engine_cpu)
    return stub


def _model():
    from transformers import Phi3Config
    from videogpt_b200 import LVM
    m = LVM(Phi3Config(**synth.REDUCED.phi3_kwargs()), device="cpu", materialize_pos_embed=False)
    return m.to(BF).eval()


def _mk(n_ctx, n_gen, H, W):
    d = po.frame_block_inputs(n_ctx, n_gen, H, W, True, 1)
    lat = [x.to(BF) for x in synth.synthetic_latents(n_ctx + n_gen, H, W)]
    mk = dict(input_ids=d["input_ids"], input_img_latents=lat[:n_ctx], input_image_sizes=d["input_image_sizes"],
              attention_mask=d["attention_mask"], position_ids=d["position_ids"],
              denoise_image_sizes=d["denoise_image_sizes"], time_emb_inx=d["time_emb_inx"], img_cfg_scale=1.5,
              use_img_cfg=True, use_kv_cache=False, offload_model=False, vae=None)
    return mk, lat[n_ctx:]


def test_scheduler_engine_loop_call_sequence(dry):
    from videogpt_b200 import LVMScheduler
    m = _model()
    mk, z0 = _mk(2, 2, 64, 64)
    out = LVMScheduler(num_steps=3)([x.clone() for x in z0] * 2, m.frame_block_forward_with_cfg, mk,
                                    use_kv_cache=False, prediction_type="x1")
    assert len(out) == 4 and out[0].shape == (1, 4, 8, 8)
    L = 2
    e = m._engine
    assert dry.calls.count("mask_from_codes") == 2                      # mask validated once per row
    assert dry.calls.count("cfg_euler") == 0 and dry.calls.count("euler_fused") == 3     # the update rides in the final kernel
    assert dry.calls.count("final_layer") == 3
    assert dry.calls.count("attention") == (L - 1) + 3 * L              # prefill skips the last layer's attention
    assert dry.calls.count("gemm") == 4 * (L - 1) + 1 + 3 * 4 * L
    assert dry.calls.count("rope_kv_append") == L + 3 * L
    n_predict = dry.calls.count("timestep_sinusoid")
    step_calls = [c for c in dry.calls[dry.calls.index("timestep_sinusoid"):] if c != "euler_fused"]
    assert e.launches_per_predict == len(step_calls) // n_predict       # the scheduler update is part of the final kernel
    # a second clip with the same layout reuses the plan (no new mask check), new context -> new prefill
    before = len(dry.calls)
    mk2 = dict(mk); mk2["input_img_latents"] = [x.clone() for x in mk["input_img_latents"]]
    LVMScheduler(num_steps=1)([x.clone() for x in z0] * 2, m.frame_block_forward_with_cfg, mk2,
                              use_kv_cache=False, prediction_type="v")
    new = dry.calls[before:]
    assert "mask_from_codes" not in new and new.count("rope_kv_append") == L + L


def test_callback_seam_and_generic_scheduler(dry, monkeypatch):
    from videogpt_b200 import LVMScheduler
    m = _model()
    mk, z0 = _mk(3, 2, 64, 96)
    t = torch.full((4,), 0.25)
    pred, cache = m.frame_block_forward_with_cfg([x.clone() for x in z0] * 2, t, past_key_values=None,
                                                 prediction_type="v", **mk)
    assert cache is None and len(pred) == 4 and pred[0].shape == (1, 4, 8, 12) and "cfg_combine" in dry.calls
    # generic loop: any callable with the reference's callback signature
    monkeypatch.setattr(torch.Tensor, "is_cuda", property(lambda self: True))
    calls = []

    def func(z, timesteps, past_key_values=None, prediction_type="v", **kw):
        calls.append(float(timesteps[0]))
        return [torch.zeros_like(x) for x in z], None

    out = LVMScheduler(num_steps=4)([x.clone() for x in z0] * 2, func, mk, prediction_type="x1")
    assert calls == [0.0, 0.25, 0.5, 0.75] and len(out) == 4


def test_pipeline_latent_entry_point(dry):
    from videogpt_b200 import LVMPipeline, LVMProcessor
    m = _model()
    pipe = LVMPipeline(None, m, LVMProcessor(synth.SingleIdTagTokenizer()), device=torch.device("cpu"))
    lat = synth.synthetic_latents(6, 64, 64)
    out = pipe.next_clip_latents(lat[:4], 2, num_inference_steps=2, img_guidance_scale=1.5, initial_noise=lat[4:])
    assert len(out) == 2 and out[0].shape == (1, 4, 8, 8)
    out = pipe.next_clip_latents(lat[:4], 2, num_inference_steps=2, img_guidance_scale=1.0, initial_noise=lat[4:])
    assert len(out) == 2                      # guidance off: single branch, all generated frames returned


def test_single_frame_path_dryrun(dry):
    from videogpt_b200 import LVMScheduler
    m = _model()
    d = po.single_frame_inputs(2, 64, 64, True, 1)
    lat = [x.to(BF) for x in synth.synthetic_latents(3, 64, 64)]
    mk = dict(input_ids=d["input_ids"], input_img_latents=lat[:2], input_image_sizes=d["input_image_sizes"],
              attention_mask=d["attention_mask"], position_ids=d["position_ids"], img_cfg_scale=1.5,
              use_img_cfg=True, use_kv_cache=False, offload_model=False)
    z = torch.cat([lat[2], lat[2]], 0)
    out = LVMScheduler(num_steps=2)(z, m.forward_with_cfg, mk, prediction_type="v")
    assert out.shape == (2, 4, 8, 8)
    pred, cache = m.forward_with_cfg(z, torch.full((2,), 0.5), past_key_values=None, prediction_type="v", **mk)
    assert cache is None and pred.shape == (2, 4, 8, 8)


# ------------------------------------------------------------------------------------------------
# sequence-parallel host logic, dry (stub kernels, CPU): every rank of a group must enqueue the SAME
# number of cross-GPU barriers in the same places, whatever the geometry / world size -- a mismatch
# would dead-lock real ranks.  (BASELINE configs[1], [2], [4] geometries at 2 / 4 / 8 ranks.)
# ------------------------------------------------------------------------------------------------
class StubOpsSP(StubOps):
    def rope_kv_append_peers(self, qkv, row_pos, row_slot, table, k_ptrs, v_ptrs, n_pools, heads, head_dim):
        self._log("rope_kv_append_peers")
        assert row_pos.numel() == row_slot.numel() == qkv.shape[0] and len(k_ptrs) == len(v_ptrs) == n_pools

    def final_layer_rows(self, hidden, kind, a, b, mod, w, bias, pred_ptrs, n_preds, lat_h, lat_w, norm_weight=None, rms_eps=0.0):
        self._log("final_layer_rows")
        assert kind.numel() == hidden.shape[0] and len(pred_ptrs) == n_preds


@pytest.mark.parametrize("geom", [(4, 4, 256, 256), (32, 4, 256, 256), (4, 4, 512, 512), (1, 1, 64, 64)])
@pytest.mark.parametrize("world", [2, 4, 8])
def test_sequence_parallel_ranks_agree_on_sync_points(monkeypatch, geom, world):
    from videogpt_b200 import engine as eng, peer
    stub = StubOpsSP()
    monkeypatch.setattr(eng, "ops", stub)
    monkeypatch.setattr(torch.cuda, "is_available", lambda: True)
    n_ctx, n_gen, H, W = geom
    dims = synth.REDUCED
    d = po.frame_block_inputs(n_ctx, n_gen, H, W, True, 1)
    specs, n_lat, n_ctx_lat = eng.frame_block_specs(d["input_ids"], d["position_ids"], d["input_image_sizes"],
                                                   d["denoise_image_sizes"], d["time_emb_inx"])
    sd = {k: v.to(BF) for k, v in synth.init_state_dict(dims, seed=0, with_pos_embed=False).items()}
    w = eng.EngineWeights(sd, dims.num_hidden_layers, "cpu")
    members = peer.LocalPeerGroup.create(world, "cpu")
    traces = []
    for r, m in enumerate(members):
        e = eng.NextClipEngine(w, dims.hidden_size, dims.intermediate_size, dims.num_hidden_layers,
                               dims.num_attention_heads, dims.rms_norm_eps, dims.rope_theta, "cpu",
                               dims.pos_embed_max_size, 2, use_cuda_graph=False, peers=m)
        e.set_plan(eng.build_plan(specs, n_lat, n_ctx_lat, H // 8, W // 8, "cpu", shard=(r, world)))
        assert e.kv.shape[2] == e.plan.total_pages and e.pred.shape == e.z.shape
        pre = list(e.prefill_steps(None))
        step = list(e.predict_steps())
        traces.append((pre, step, e.plan.prefix.rows, e.plan.step.rows))
    L = dims.num_hidden_layers
    for pre, step, _, _ in traces:
        assert pre == (["kv"] * L if n_ctx else [])                 # one barrier per layer's K/V append
        assert step == ["kv"] * L + ["pred"]                        # + one for the prediction
    assert sum(t[2] for t in traces) == sum(sp.n_prefix for sp in specs)
    assert sum(t[3] for t in traces) == sum(sp.n_active for sp in specs)


@pytest.mark.parametrize("geom", [(4, 4, 256, 256), (32, 4, 256, 256), (1, 1, 64, 64)])
def test_cfg_branch_pair_ranks_agree_on_sync_points(monkeypatch, geom):
    """partition="sequences" on two ranks: no barrier in the prefill (K/V never travel: the local append kernel, not
    the peer one), two per step -- one in front (a peer's next prediction must not overwrite one this rank has not
    consumed) and one behind the prediction stores -- on BOTH ranks although rank 1 has no prefix rows at all; the
    scheduler update is enqueued behind the second barrier when the sampler's fused loop asks for it."""
    from videogpt_b200 import engine as eng, peer
    stub = StubOpsSP()
    monkeypatch.setattr(eng, "ops", stub)
    monkeypatch.setattr(torch.cuda, "is_available", lambda: True)
    n_ctx, n_gen, H, W = geom
    dims = synth.REDUCED
    d = po.frame_block_inputs(n_ctx, n_gen, H, W, True, 1)
    specs, n_lat, n_ctx_lat = eng.frame_block_specs(d["input_ids"], d["position_ids"], d["input_image_sizes"],
                                                   d["denoise_image_sizes"], d["time_emb_inx"])
    sd = {k: v.to(BF) for k, v in synth.init_state_dict(dims, seed=0, with_pos_embed=False).items()}
    w = eng.EngineWeights(sd, dims.num_hidden_layers, "cpu")
    L = dims.num_hidden_layers
    for r, m in enumerate(peer.LocalPeerGroup.create(2, "cpu")):
        e = eng.NextClipEngine(w, dims.hidden_size, dims.intermediate_size, dims.num_hidden_layers,
                               dims.num_attention_heads, dims.rms_norm_eps, dims.rope_theta, "cpu",
                               dims.pos_embed_max_size, 2, use_cuda_graph=False, peers=m)
        e.set_plan(eng.build_plan(specs, n_lat, n_ctx_lat, H // 8, W // 8, "cpu", shard=(r, 2), partition="sequences"))
        stub.calls.clear()
        assert list(e.prefill_steps(None)) == []
        assert (e.plan.prefix.rows, e.plan.step.rows) == ((specs[0].n_prefix, specs[0].n_active) if r == 0 else (0, specs[1].n_active))
        assert stub.calls.count("rope_kv_append") == (L if r == 0 else 0) and "rope_kv_append_peers" not in stub.calls
        stub.calls.clear()
        e.euler_mode = (True, True)
        assert list(e.predict_steps()) == ["start", "pred"]
        assert stub.calls.count("rope_kv_append") == L and "rope_kv_append_peers" not in stub.calls
        assert stub.calls.count("final_layer_rows") == 1 and stub.calls[-1] == "cfg_euler"
        m.lockstep = False
        assert e.launches_per_predict == 6 + 1 + 8 * L + 1 + 2


def test_plan_cache_keeps_its_key_objects_alive(dry):
    """The per-clip cache key identifies the conditioning tensors by object; the model must hold
    them, otherwise the allocator hands the freed address to the NEXT clip's context and a stale
    prefill (old context K/V) would be reused silently."""
    import gc
    import weakref
    from videogpt_b200 import LVMScheduler
    m = _model()
    mk, z0 = _mk(2, 2, 64, 64)
    probe = weakref.ref(mk["input_img_latents"][0])
    LVMScheduler(num_steps=1)([x.clone() for x in z0] * 2, m.frame_block_forward_with_cfg, mk, prediction_type="x1")
    n_prefill = dry.calls.count("rope_kv_append")
    ids, pos = mk["input_ids"], mk["position_ids"]
    del mk
    gc.collect()
    assert probe() is not None                       # still referenced by the model's cache
    # a new clip with NEW tensor objects (same values, same layout) must prefill again
    mk2, _ = _mk(2, 2, 64, 64)
    mk2["input_ids"], mk2["position_ids"] = ids, pos
    LVMScheduler(num_steps=1)([x.clone() for x in z0] * 2, m.frame_block_forward_with_cfg, mk2, prediction_type="x1")
    L = 2
    assert dry.calls.count("rope_kv_append") == n_prefill + L + L      # prefill (L) + one step (L)
    # the same objects again (every Euler step of one clip through the callback seam): no new prefill
    before = dry.calls.count("rope_kv_append")
    m.frame_block_forward_with_cfg([x.clone() for x in z0] * 2, torch.full((4,), 0.5), past_key_values=None,
                                   prediction_type="x1", **mk2)
    assert dry.calls.count("rope_kv_append") == before + L
