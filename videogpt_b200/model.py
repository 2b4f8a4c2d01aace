"""``LVM``: drop-in for the reference's inference model wrapper (``LVM/model.py:157-566``).

Same constructor, state-dict names, ``from_pretrained`` and the two scheduler callbacks
``frame_block_forward_with_cfg`` / ``forward_with_cfg`` (seam S2 of SURVEY.md 8(b)); the
arithmetic runs in ``NextClipEngine`` on hand-written sm_100a kernels.  There is no CPU
path: calling a forward without CUDA raises.

Differences from the reference that a caller can observe:
* ``past_key_values`` is accepted and ignored (the reference's LVM path never reuses it:
  ``LVM/scheduler.py:174``); the returned cache is ``None``.  Context K/V are cached
  internally across the calls of one clip, keyed on the identity of the conditioning inputs.
* ``attention_mask`` is validated once against the closed form and then not read.
* compute dtype is bf16 (``pipeline`` default ``dtype=torch.bfloat16``, pipeline.py:361).
"""
from __future__ import annotations

import os
from typing import Dict, Optional

import numpy as np
import torch
import torch.nn as nn

from . import engine as eng
from . import ops
from .parallel_states import hccl_info
from .synth import BackboneDims, sincos_pos_embed_table


def _cfg_get(cfg, name, default=None):
    v = getattr(cfg, name, None)
    if v is None and name == "rope_theta":
        for attr in ("rope_parameters", "rope_scaling"):
            rp = getattr(cfg, attr, None)
            if isinstance(rp, dict) and rp.get("rope_theta") is not None:
                v = rp["rope_theta"]
                break
    return default if v is None else v


class _Linear(nn.Module):
    def __init__(self, n_in, n_out, bias=True):
        super().__init__()
        self.weight = nn.Parameter(torch.empty(n_out, n_in))
        self.bias = nn.Parameter(torch.zeros(n_out)) if bias else None


class _Seq(nn.Module):
    """Holds children under the numeric names ``nn.Sequential`` would give them."""
    def __init__(self, mods: Dict[str, nn.Module]):
        super().__init__()
        for k, m in mods.items():
            self.add_module(k, m)


def _k(t):
    """Parameter / input as the kernels take it: detached, contiguous, activation dtype."""
    return t.detach().to(eng.ACT_DTYPE).contiguous()


class TimestepEmbedder(nn.Module):          # LVM/model.py:26-63
    def __init__(self, hidden_size, frequency_embedding_size=256):
        super().__init__()
        self.mlp = _Seq({"0": _Linear(frequency_embedding_size, hidden_size), "2": _Linear(hidden_size, hidden_size)})
        self.frequency_embedding_size = frequency_embedding_size

    @torch.no_grad()
    def forward(self, t, dtype=None):
        """``TimestepEmbedder.forward`` (model.py:60-63): sinusoid ([cos | sin], fp32 -> model dtype) ->
        Linear -> SiLU -> Linear, on the kernels of the hot path.  ``t``: ``[n]`` on the model's device."""
        import math
        half = self.frequency_embedding_size // 2
        l0, l2 = getattr(self.mlp, "0"), getattr(self.mlp, "2")
        dev = l0.weight.device
        freqs = torch.exp(-math.log(10000.0) * torch.arange(half, dtype=torch.float32) / half).to(dev)
        x = torch.empty(t.numel(), 2 * half, device=dev, dtype=eng.ACT_DTYPE)
        ops.timestep_sinusoid(t.to(dev, torch.float32).contiguous(), freqs, x)
        hdn = ops.linear_small(x, _k(l0.weight), _k(l0.bias), post_silu=True)
        return ops.linear_small(hdn, _k(l2.weight), _k(l2.bias))


class FinalLayer(nn.Module):                # LVM/model.py:66-83
    def __init__(self, hidden_size, patch_size, out_channels):
        super().__init__()
        self.linear = _Linear(hidden_size, patch_size * patch_size * out_channels)
        self.adaLN_modulation = _Seq({"1": _Linear(hidden_size, 2 * hidden_size)})
        self.patch_size, self.out_channels = patch_size, out_channels

    @torch.no_grad()
    def forward(self, x, c):
        """``FinalLayer.forward`` (model.py:79-83): ``x [B, T, hidden]``, ``c [B, hidden]`` ->
        ``[B, T, p*p*C]`` (feature order (p, q, c)).  Runs the fused final-layer kernel -- which
        scatters into latent layout -- on a one-patch-row latent of 2 x 2T pixels and reads it back
        in token order (indexing only)."""
        b, t, hs = x.shape
        p, ch = self.patch_size, self.out_channels
        ada = getattr(self.adaLN_modulation, "1")
        mod = ops.linear_small(_k(c), _k(ada.weight), _k(ada.bias), pre_silu=True)
        pred = torch.empty(b, ch, p, p * t, device=x.device, dtype=eng.ACT_DTYPE)
        row0 = (torch.arange(b, dtype=torch.int32, device=x.device) * t).contiguous()
        ops.final_layer(_k(x).reshape(b * t, hs), row0, mod, _k(self.linear.weight), _k(self.linear.bias), pred)
        return pred.reshape(b, ch, p, t, p).permute(0, 3, 2, 4, 1).reshape(b, t, p * p * ch)


class PatchEmbedMR(nn.Module):              # LVM/model.py:138-154
    def __init__(self, patch_size=2, in_chans=4, embed_dim=768, bias=True):
        super().__init__()
        self.proj = nn.Module()
        self.proj.weight = nn.Parameter(torch.empty(embed_dim, in_chans, patch_size, patch_size))
        self.proj.bias = nn.Parameter(torch.zeros(embed_dim))
        self.patch_size = patch_size

    @torch.no_grad()
    def forward(self, latent, pos_rows=None):
        """``PatchEmbedMR.forward`` (model.py:149-153): ``[B, 4, h, w]`` -> ``[B, tokens, hidden]`` (conv with
        kernel = stride = patch, flattened row-major), through the assembly kernel; ``pos_rows``
        ``[tokens, hidden]`` is added in the same pass when given (the model always adds one)."""
        b, _, h, w = latent.shape
        n_tok = (h // self.patch_size) * (w // self.patch_size)
        dev, hs = latent.device, self.proj.weight.shape[0]
        pos = torch.zeros(n_tok, hs, device=dev, dtype=eng.ACT_DTYPE) if pos_rows is None else _k(pos_rows)
        kind = torch.full((b * n_tok,), ops.ROW_NOISY_PATCH, dtype=torch.int32, device=dev)
        a = torch.arange(b, dtype=torch.int32, device=dev).repeat_interleave(n_tok)
        t = torch.arange(n_tok, dtype=torch.int32, device=dev).repeat(b)
        out = torch.empty(b * n_tok, hs, device=dev, dtype=eng.ACT_DTYPE)
        z = _k(latent)
        wt, bs = _k(self.proj.weight), _k(self.proj.bias)
        ops.embed_assemble(out, kind, a, t, wt.reshape(hs, -1), None, z, z, h, w, wt, bs, wt, bs, pos)
        return out.reshape(b, n_tok, hs)


class _Norm(nn.Module):
    def __init__(self, n):
        super().__init__()
        self.weight = nn.Parameter(torch.ones(n))


class _Attn(nn.Module):
    def __init__(self, h):
        super().__init__()
        self.qkv_proj = _Linear(h, 3 * h, bias=False)
        self.o_proj = _Linear(h, h, bias=False)


class _MLP(nn.Module):
    def __init__(self, h, i):
        super().__init__()
        self.gate_up_proj = _Linear(h, 2 * i, bias=False)
        self.down_proj = _Linear(i, h, bias=False)


class _Layer(nn.Module):
    def __init__(self, h, i):
        super().__init__()
        self.self_attn, self.mlp = _Attn(h), _MLP(h, i)
        self.input_layernorm, self.post_attention_layernorm = _Norm(h), _Norm(h)


class Phi3Backbone(nn.Module):
    """Parameter container with the names of ``Phi3Transformer`` (OmniGen/transformer.py:35)."""
    def __init__(self, config):
        super().__init__()
        self.config = config
        h, i = config.hidden_size, config.intermediate_size
        self.embed_tokens = nn.Module()
        self.embed_tokens.weight = nn.Parameter(torch.empty(config.vocab_size, h))
        self.layers = nn.ModuleList([_Layer(h, i) for _ in range(config.num_hidden_layers)])
        self.norm = _Norm(h)


class LVM(nn.Module):
    def __init__(self, transformer_config, patch_size=2, in_channels=4, pe_interpolation: float = 1.0,
                 pos_embed_max_size: int = 192, device=None, materialize_pos_embed: bool = True):
        super().__init__()
        self.in_channels = self.out_channels = in_channels
        self.patch_size = patch_size
        self.pos_embed_max_size = pos_embed_max_size
        self.pe_interpolation = pe_interpolation
        hidden = transformer_config.hidden_size
        self.hidden_size = hidden
        nh = transformer_config.num_attention_heads
        nkv = _cfg_get(transformer_config, "num_key_value_heads", nh)
        if nkv != nh:
            raise ValueError("grouped-query attention is not supported (the reference model uses 32 = 32 heads)")
        rs = _cfg_get(transformer_config, "rope_scaling")
        if isinstance(rs, dict):      # transformers >= 5 folds rope_theta / rope_type into this dict
            rs = None if rs.get("rope_type", rs.get("type", "default")) == "default" else rs
        if rs is not None:
            raise ValueError("rope_scaling is not supported (the reference uses Phi3Config defaults)")
        if patch_size != 2 or in_channels != 4:
            raise ValueError("only patch_size=2, in_channels=4 (SDXL-VAE latents) are supported")
        with torch.device(device if device is not None else ("cuda" if torch.cuda.is_available() else "cpu")):
            self.x_embedder = PatchEmbedMR(patch_size, in_channels, hidden)
            self.input_x_embedder = PatchEmbedMR(patch_size, in_channels, hidden)
            self.time_token = TimestepEmbedder(hidden)
            self.t_embedder = TimestepEmbedder(hidden)
            self.final_layer = FinalLayer(hidden, patch_size, self.out_channels)
            self.llm = Phi3Backbone(transformer_config)
        self.llm.config.use_cache = False
        if materialize_pos_embed:
            pe = sincos_pos_embed_table(hidden, pos_embed_max_size, 64, pe_interpolation)
            self.register_buffer("pos_embed", pe.to(self.x_embedder.proj.weight.device), persistent=True)
        else:
            self.pos_embed = None
        self.initialize_weights()
        self._engine: Optional[eng.NextClipEngine] = None
        self._engine_key = None
        self._plan_key = None
        self._layout_key = None
        self._plan_refs = None
        self._peers = None
        self.use_cuda_graph = True

    # ---- construction / weights ---------------------------------------------------------------
    def initialize_weights(self):
        """Same distributions as ``LVM.initialize_weights`` (model.py:213-244) and the HF Phi-3
        init (N(0, initializer_range)); final layer zero like the reference."""
        def xavier(w):
            nn.init.xavier_uniform_(w.view(w.shape[0], -1))
        std = _cfg_get(self.llm.config, "initializer_range", 0.02)
        with torch.no_grad():
            xavier(self.x_embedder.proj.weight); xavier(self.input_x_embedder.proj.weight)
            for e in (self.time_token, self.t_embedder):
                nn.init.normal_(getattr(e.mlp, "0").weight, std=0.02)
                nn.init.normal_(getattr(e.mlp, "2").weight, std=0.02)
            for lin in (self.final_layer.linear, getattr(self.final_layer.adaLN_modulation, "1")):
                nn.init.zeros_(lin.weight); nn.init.zeros_(lin.bias)
            nn.init.normal_(self.llm.embed_tokens.weight, std=std)
            for layer in self.llm.layers:
                for lin in (layer.self_attn.qkv_proj, layer.self_attn.o_proj, layer.mlp.gate_up_proj,
                            layer.mlp.down_proj):
                    nn.init.normal_(lin.weight, std=std)

    @classmethod
    def from_pretrained(cls, model_name, load_llm_ckpt=True):
        """``LVM.from_pretrained`` (model.py:195-211): config + ``model.safetensors`` / ``model.pt``."""
        from transformers import Phi3Config
        if not os.path.exists(model_name):
            from huggingface_hub import snapshot_download
            model_name = snapshot_download(repo_id=model_name, cache_dir=os.getenv("HF_HUB_CACHE"),
                                           ignore_patterns=["flax_model.msgpack", "rust_model.ot", "tf_model.h5"])
        model = cls(Phi3Config.from_pretrained(model_name))
        if load_llm_ckpt:
            st = os.path.join(model_name, "model.safetensors")
            if os.path.exists(st):
                from safetensors.torch import load_file
                ckpt = load_file(st)
            else:
                ckpt = torch.load(os.path.join(model_name, "model.pt"), map_location="cpu")
            model.load_state_dict(ckpt)
        return model

    def load_state_dict(self, state_dict, strict: bool = True, **kw):
        self._engine = None
        return super().load_state_dict(state_dict, strict=strict, **kw)

    def _apply(self, fn, *a, **k):
        ref = self.final_layer.linear.weight if hasattr(self, "final_layer") else None
        before = None if ref is None else (ref.device, ref.dtype, ref.data_ptr())
        out = super()._apply(fn, *a, **k)
        ref = self.final_layer.linear.weight if hasattr(self, "final_layer") else None
        if ref is None or before != (ref.device, ref.dtype, ref.data_ptr()):
            self._engine = None           # parameters moved / were cast: rebuild the engine lazily
        return out

    def dims(self) -> BackboneDims:
        c = self.llm.config
        return BackboneDims(hidden_size=c.hidden_size, intermediate_size=c.intermediate_size,
                            num_hidden_layers=c.num_hidden_layers, num_attention_heads=c.num_attention_heads,
                            rms_norm_eps=_cfg_get(c, "rms_norm_eps", 1e-5),
                            rope_theta=float(_cfg_get(c, "rope_theta", 10000.0)),
                            vocab_size=c.vocab_size, pos_embed_max_size=self.pos_embed_max_size)

    # ---- the reference model's helper methods (LVM/model.py:255-327) -----------------------------------
    def unpatchify(self, x, h, w):
        """``LVM.unpatchify`` (model.py:255-265): ``[N, T, p*p*C]`` -> ``[N, C, h, w]`` (pure data movement;
        on the hot path the final-layer kernel scatters straight into this layout)."""
        c, p = self.out_channels, self.patch_size
        x = x.reshape(x.shape[0], h // p, w // p, p, p, c)
        return torch.einsum("nhwpqc->nchpwq", x).reshape(x.shape[0], c, h, w)

    def cropped_pos_embed(self, height, width):
        """``LVM.cropped_pos_embed`` (model.py:268-289): centre crop of the position table for a latent of
        ``height x width``, ``[1, tokens, hidden]`` in the model's dtype (a gather: exact)."""
        if self.pos_embed_max_size is None:
            raise ValueError("`pos_embed_max_size` must be set for cropping.")
        hh, ww = height // self.patch_size, width // self.patch_size
        if hh > self.pos_embed_max_size:
            raise ValueError(f"Height ({hh}) cannot be greater than `pos_embed_max_size`: {self.pos_embed_max_size}.")
        if ww > self.pos_embed_max_size:
            raise ValueError(f"Width ({ww}) cannot be greater than `pos_embed_max_size`: {self.pos_embed_max_size}.")
        ref = self.x_embedder.proj.weight
        if self.pos_embed is not None:
            top, left = (self.pos_embed_max_size - hh) // 2, (self.pos_embed_max_size - ww) // 2
            pe = self.pos_embed.reshape(1, self.pos_embed_max_size, self.pos_embed_max_size, -1)
            return pe[:, top:top + hh, left:left + ww, :].reshape(1, hh * ww, -1)
        from .synth import cropped_pos_embed_rows
        return cropped_pos_embed_rows(self.hidden_size, height, width, self.patch_size,
                                      self.pos_embed_max_size).to(ref.device, ref.dtype)[None]

    @torch.no_grad()
    def patch_multiple_resolutions(self, latents, padding_latent=None, is_input_images: bool = False):
        """``LVM.patch_multiple_resolutions`` (model.py:292-327): patch embedding (``x_embedder`` /
        ``input_x_embedder``) + cropped position embedding of a latent ``[B,4,h,w]`` or of a list of them,
        through the assembly kernel (``vgpt_embed_assemble``).  Returns ``(latents, num_tokens, shapes)``
        like the reference."""
        def embed(lat):
            h, w = lat.shape[-2:]
            emb = (self.input_x_embedder if is_input_images else self.x_embedder)(lat, self.cropped_pos_embed(h, w)[0])
            return emb, emb.shape[1], [h, w]

        if not isinstance(latents, list):
            return embed(latents)
        return_list = padding_latent is None
        pads = [None] * len(latents) if return_list else padding_latent
        patched, num_tokens, shapes = [], [], []
        for lat, pad in zip(latents, pads):
            e, n_tok, shape = embed(lat)
            if pad is not None:
                e = torch.cat([e, pad], dim=-2)
            patched.append(e)
            num_tokens.append(n_tok)
            shapes.append(shape)
        return (patched if return_list else torch.cat(patched, dim=0)), num_tokens, shapes

    # ---- engine ---------------------------------------------------------------------------------
    def engine(self) -> eng.NextClipEngine:
        dev = self.x_embedder.proj.weight.device
        if dev.type != "cuda":
            raise RuntimeError("videogpt_b200.LVM runs on CUDA (sm_100a) only: move the model with "
                               ".to('cuda'); there is no CPU fallback")
        peers = self.sequence_parallel_peers()
        key = (dev, tuple(p._version for p in self.parameters()), self.use_cuda_graph, id(peers))
        if self._engine is None or self._engine_key != key:
            d = self.dims()
            sd = {k: v for k, v in self.state_dict().items()}
            w = eng.EngineWeights(sd, d.num_hidden_layers, dev)
            self._engine = eng.NextClipEngine(w, d.hidden_size, d.intermediate_size, d.num_hidden_layers,
                                              d.num_attention_heads, d.rms_norm_eps, d.rope_theta, dev,
                                              self.pos_embed_max_size, self.patch_size, self.use_cuda_graph,
                                              peers=peers)
            self._engine_key, self._plan_key, self._layout_key = key, None, None
        return self._engine

    def sequence_parallel_peers(self):
        """Peer group of this rank when ``initialize_sequence_parallel_state(P > 1)`` ran (the
        reference's switch for sequence parallelism, ``LVM/model.py:459``): the ranks of
        ``hccl_info.group`` then share every video -- each computes a contiguous chunk of the rows
        of every sequence and stores its K/V and predictions into all peers (``peer.py``); or, with
        ``hccl_info.partition == "sequences"``, whole sequences (one CFG branch per rank of a pair)
        and only the predictions travel."""
        if hccl_info.world_size in (0, 1, None):
            return None
        if self._peers is None:
            import torch.distributed as dist
            from . import peer
            ranks = dist.get_process_group_ranks(hccl_info.group) if hccl_info.group is not None else \
                list(range(dist.get_world_size()))
            self._peers = peer.PeerGroup(ranks, group=hccl_info.group, host_group=getattr(hccl_info, "host_group", None))
        return self._peers

    def invalidate_plan_cache(self):
        """Forget which conditioning inputs the engine's current plan belongs to (callers that set
        engine plans themselves -- ``rollout.LatentRollout`` -- call this so that the next
        ``prepare_*`` rebuilds instead of trusting a plan it did not make)."""
        self._plan_key = self._layout_key = self._plan_refs = None

    def _shard(self):
        """``shard=`` / ``partition=`` of ``engine.build_plan`` for this rank (nothing on a single GPU)."""
        e = self._engine
        if e is None or e.peers is None:
            return {}
        return dict(shard=(e.peers.rank, e.peers.world), partition=getattr(hccl_info, "partition", "rows"))

    @staticmethod
    def _identity(*objs):
        """Identity of (nested) conditioning inputs: tensors by OBJECT (id + storage + version).

        Valid only while the objects are alive -- an address or id freed by one clip is routinely
        handed to the next clip's tensors by the allocator -- so whoever stores an identity as a
        cache key must also keep the objects (``_plan_refs``).  A hit then means "the very same,
        unmodified tensors", which is also the same decision on every rank of a sequence-parallel
        group (it does not depend on any rank's allocator state)."""
        def freeze(o):
            if torch.is_tensor(o):
                return (id(o), o.data_ptr(), tuple(o.shape), o._version)
            if isinstance(o, (list, tuple)):
                return tuple(freeze(x) for x in o)
            if isinstance(o, dict):
                return tuple((k, freeze(v)) for k, v in o.items())
            return o
        return tuple(freeze(o) for o in objs)

    def prepare_frame_block(self, input_ids, input_img_latents, input_image_sizes, attention_mask,
                            position_ids, denoise_image_sizes, time_emb_inx, lat_h, lat_w,
                            check_mask: bool = True):
        """Build (or reuse) the plan for these conditioning inputs and prefill the context K/V.

        Three cache levels, cheapest first: (1) same tensor objects as the previous call (every
        Euler step of one clip) -> nothing to do; (2) same token layout (next clip of the same
        geometry) -> keep plan, workspaces and the captured CUDA graph, redo only the prefill;
        (3) new layout -> new plan."""
        e = self.engine()             # sequence parallel (model.py:459-464) when hccl_info says so
        ident = self._identity(input_ids, position_ids, input_img_latents, input_image_sizes,
                               denoise_image_sizes, time_emb_inx, attention_mask, lat_h, lat_w)
        if ident == self._plan_key and e.plan is not None and e.prefilled:
            return e
        ids_host, pos_host = input_ids.cpu(), position_ids.cpu()
        layout = (ids_host.numpy().tobytes(), pos_host.numpy().tobytes(), tuple(ids_host.shape),
                  self._identity(input_image_sizes, denoise_image_sizes, time_emb_inx), lat_h, lat_w,
                  hccl_info.partition)
        n_ctx = sum(len(v) for v in input_image_sizes.values())
        if n_ctx != len(input_img_latents or []):
            raise AssertionError("number of context latents does not match input_image_sizes")   # model.py:454
        if layout != self._layout_key or e.plan is None:
            specs, n_lat, n_ctx = eng.frame_block_specs(ids_host, pos_host, input_image_sizes,
                                                       denoise_image_sizes, time_emb_inx)
            if check_mask and attention_mask is not None:
                self._check_mask(attention_mask, specs, e.device)
            e.set_plan(eng.build_plan(specs, n_lat, n_ctx, lat_h, lat_w, e.device, **self._shard()))
            self._layout_key = layout
        ctx = torch.cat([x.reshape(1, 4, lat_h, lat_w) for x in input_img_latents], 0) if n_ctx else None
        e.prefill(ctx)
        self._plan_key = ident
        self._plan_refs = (input_ids, position_ids, list(input_img_latents or []), attention_mask)   # keep the key's objects alive
        return e

    @staticmethod
    def _check_mask(attention_mask, specs, device):
        """The kernels derive the mask from token codes; reject any other mask loudly
        (``Phi3Transformer.forward`` raises on unusable masks too: OmniGen/transformer.py:151)."""
        if attention_mask.dim() != 3:
            raise Exception("attention_mask parameter was unavailable or invalid")
        L = attention_mask.shape[-1]
        for b, sp in enumerate(specs):
            T = sp.n_prefix + sp.n_active
            pad = L - T
            qc = torch.from_numpy(np.concatenate([np.full(pad, eng.INT_MAX), sp.codes]).astype(np.int32)).to(device)
            kc = torch.from_numpy(np.concatenate([np.full(pad, eng.INT_MAX - 1), sp.codes]).astype(np.int32)).to(device)
            want = ops.mask_from_codes(qc, kc)
            got = attention_mask[b].to(device=device)
            if not torch.equal(want.bool(), got.bool()):
                raise ValueError("attention_mask is not the frame-block mask of these index dicts; "
                                 "videogpt_b200 only implements the reference's closed-form masks")

    # ---- seam S2: scheduler callbacks -----------------------------------------------------------
    @torch.no_grad()
    def frame_block_forward(self, x, timestep, input_ids, input_img_latents, input_image_sizes,
                            attention_mask, position_ids, denoise_image_sizes, time_emb_inx,
                            padding_latent=None, past_key_values=None, return_past_key_values=True,
                            offload_model: bool = False, vae=None, input_output_return=False):
        """``LVM.frame_block_forward`` (model.py:399-501).  ``x``: list of ``[1,4,h,w]`` latents."""
        if input_output_return or padding_latent is not None:
            raise NotImplementedError("training-only arguments are out of scope")
        assert input_ids is not None, "input_ids is None"
        lat_h, lat_w = x[0].shape[-2:]
        e = self.prepare_frame_block(input_ids, input_img_latents, input_image_sizes, attention_mask,
                                     position_ids, denoise_image_sizes, time_emb_inx, lat_h, lat_w)
        n = e.plan.n_latents
        assert len(x) == n and timestep.numel() == n                                  # model.py:454
        e.z.copy_(torch.cat([t.reshape(1, 4, lat_h, lat_w) for t in x], 0))
        e.t.copy_(timestep.to(device=e.device, dtype=torch.float32))
        pred = e.predict()
        latents = [pred[i:i + 1].clone() for i in range(n)]
        return (latents, None) if return_past_key_values else latents

    @torch.no_grad()
    def frame_block_forward_with_cfg(self, x, timestep, input_ids, input_img_latents, input_image_sizes,
                                     attention_mask, position_ids, denoise_image_sizes, time_emb_inx,
                                     use_img_cfg, img_cfg_scale, past_key_values, use_kv_cache,
                                     offload_model, vae, prediction_type: str = "v"):
        """``LVM.frame_block_forward_with_cfg`` (model.py:518-566)."""
        self.llm.config.use_cache = use_kv_cache
        out, _ = self.frame_block_forward(x, timestep, input_ids, input_img_latents, input_image_sizes,
                                          attention_mask, position_ids, denoise_image_sizes, time_emb_inx)
        if use_img_cfg and prediction_type == "v":
            e = self._engine
            assert len(out) % 2 == 0
            ops.cfg_combine(e.pred, img_cfg_scale)
            half = len(out) // 2
            cond = [e.pred[i:i + 1].clone() for i in range(half)]
            out = cond + cond
        return out, None

    # ---- pipeline.__call__ path (one output frame, batched tensor x) --------------------------------
    def prepare_single_frame(self, input_ids, input_img_latents, input_image_sizes, attention_mask,
                             position_ids, lat_h, lat_w, check_mask: bool = True):
        e = self.engine()
        ident = self._identity("single", input_ids, position_ids, input_img_latents, input_image_sizes,
                               attention_mask, lat_h, lat_w)
        if ident == self._plan_key and e.plan is not None and e.prefilled:
            return e
        ids_host, pos_host = input_ids.cpu(), position_ids.cpu()
        layout = ("single", ids_host.numpy().tobytes(), pos_host.numpy().tobytes(), tuple(ids_host.shape),
                  self._identity(input_image_sizes), lat_h, lat_w, hccl_info.partition)
        n_tok = (lat_h // self.patch_size) * (lat_w // self.patch_size)
        n_ctx = sum(len(v) for v in input_image_sizes.values())
        assert n_ctx == len(input_img_latents or [])                                   # model.py:358
        for lat in (input_img_latents or []):
            if tuple(lat.shape[-2:]) != (lat_h, lat_w):
                raise NotImplementedError("context and output frames must have the same size "
                                          "(use_input_image_size_as_output=True)")
        if layout != self._layout_key or e.plan is None:
            specs, n_lat, n_ctx = eng.single_frame_specs(ids_host, pos_host, input_image_sizes, n_tok)
            if check_mask and attention_mask is not None:
                self._check_mask(attention_mask, specs, e.device)
            e.set_plan(eng.build_plan(specs, n_lat, n_ctx, lat_h, lat_w, e.device, **self._shard()))
            self._layout_key = layout
        ctx = torch.cat([x.reshape(1, 4, lat_h, lat_w) for x in input_img_latents], 0) if n_ctx else None
        e.prefill(ctx)
        self._plan_key = ident
        self._plan_refs = (input_ids, position_ids, list(input_img_latents or []), attention_mask)
        return e

    @torch.no_grad()
    def forward(self, x, timestep, input_ids, input_img_latents, input_image_sizes, attention_mask,
                position_ids, padding_latent=None, past_key_values=None, return_past_key_values=True,
                offload_model: bool = False):
        """``LVM.forward`` (model.py:330-397) for a batched tensor ``x`` ``[B,4,h,w]``."""
        if isinstance(x, list) or padding_latent is not None:
            raise NotImplementedError("multi-resolution lists are not supported on this path")
        lat_h, lat_w = x.shape[-2:]
        e = self.prepare_single_frame(input_ids, input_img_latents, input_image_sizes, attention_mask,
                                      position_ids, lat_h, lat_w)
        assert x.shape[0] == e.plan.n_latents and timestep.numel() == x.shape[0]
        e.z.copy_(x)
        e.t.copy_(timestep.to(device=e.device, dtype=torch.float32))
        out = e.predict().clone()
        return (out, None) if return_past_key_values else out

    @torch.no_grad()
    def forward_with_cfg(self, x, timestep, input_ids, input_img_latents, input_image_sizes, attention_mask,
                         position_ids, use_img_cfg, img_cfg_scale, past_key_values, use_kv_cache,
                         offload_model, prediction_type: str = "v"):
        """``LVM.forward_with_cfg`` (model.py:503-516)."""
        self.llm.config.use_cache = use_kv_cache
        out, _ = self.forward(x, timestep, input_ids, input_img_latents, input_image_sizes, attention_mask,
                              position_ids)
        if use_img_cfg and prediction_type == "v":
            ops.cfg_combine(self._engine.pred, img_cfg_scale)
            out = self._engine.pred.clone()
        return out, None
