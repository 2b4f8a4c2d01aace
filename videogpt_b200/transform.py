"""Operator seam S1: ``replace_attention`` (reference: ``LVM/transform/sdpa_transform.py:162-169``).

The reference patches every Phi-3 attention module of ``model.llm`` with a ``forward`` that runs
the fused qkv projection, RoPE, (Ulysses all-to-all,) SDPA with the dense additive mask and the
output projection.  This module installs a ``forward`` with the same signature and return value
that runs the B200 kernels instead: tcgen05 GEMMs for qkv / o, the fused RoPE + KV-append kernel,
and the clip-block-causal tcgen05 attention kernel.

The attention kernel does not read masks.  The additive ``[B,1,L,L]`` mask the caller passes is
converted ONCE per mask tensor into per-token codes (``allowed(q,k) <=> code_q >= code_k``) and the
conversion is verified against the mask; a mask that is not of that block-causal form (every mask
``LVM/processor.py`` builds is) raises ``ValueError``.  Sequence-parallel chunks (``S != L``) and
``past_key_value`` are not served at this seam (the engine of ``videogpt_b200.LVM`` owns caching).
"""
from __future__ import annotations

import math
from types import MethodType
from typing import Optional

import torch

from . import ops
from .ops import PAGE_TOKENS, ATTN_KV_TILE

_INT_MAX = 2 ** 31 - 1


def codes_from_mask(allowed: torch.Tensor):
    """``allowed`` bool [L, L] (row = query) -> (q_code, k_code) int32 [L] with
    ``allowed == (q_code[:, None] >= k_code[None, :])``, or ValueError."""
    L = allowed.shape[-1]
    rows = allowed.sum(-1).to(torch.int32)                       # a query's code = how many keys it sees
    big = torch.full((L, L), _INT_MAX, dtype=torch.int32, device=allowed.device)
    k_code = torch.where(allowed, rows[:, None].expand(L, L), big).amin(0)   # weakest query that sees k
    if not torch.equal(rows[:, None] >= k_code[None, :], allowed):
        raise ValueError("attention mask is not of the block-causal form code_q >= code_k; "
                         "videogpt_b200 only implements the masks LVM/processor.py builds")
    return rows.contiguous(), k_code.contiguous()


class _MaskPlan:
    def __init__(self, additive_mask: torch.Tensor):
        B, _, L, _ = additive_mask.shape
        dev = additive_mask.device
        self.B, self.L = B, L
        pages = (L + PAGE_TOKENS - 1) // PAGE_TOKENS
        self.pages = pages
        qc, kc = zip(*(codes_from_mask(additive_mask[b, 0] == 0) for b in range(B)))
        self.q_code = torch.cat(qc)
        k_code = torch.full((B, pages * PAGE_TOKENS), _INT_MAX, dtype=torch.int32, device=dev)
        k_code[:, :L] = torch.stack(kc)
        self.k_code = k_code
        tiles = k_code.view(B, pages * PAGE_TOKENS // ATTN_KV_TILE, ATTN_KV_TILE)
        self.k_tile_minmax = torch.stack([tiles.amin(-1), tiles.amax(-1)], -1).contiguous()
        self.page_table = torch.arange(B * pages, dtype=torch.int32, device=dev).view(B, pages)
        self.seqs = torch.tensor([[b * L, L, L, 0] for b in range(B)], dtype=torch.int32, device=dev)
        logical = torch.arange(L, device=dev)
        self.row_slot = (self.page_table[:, (logical // PAGE_TOKENS)].to(torch.int64) * PAGE_TOKENS
                         + logical % PAGE_TOKENS).reshape(-1).to(torch.int32)


def _mask_plan(module, attention_mask):
    key = (attention_mask.data_ptr(), tuple(attention_mask.shape), attention_mask._version)
    cache = module.__dict__.setdefault("_vgpt_mask_cache", {})
    if cache.get("key") != key:
        cache.clear()
        cache.update(key=key, plan=_MaskPlan(attention_mask))
    return cache["plan"]


def _geometry(module):
    cfg = getattr(module, "config", None)
    heads = getattr(module, "num_heads", None) or cfg.num_attention_heads
    head_dim = getattr(module, "head_dim", None) or cfg.hidden_size // heads
    theta = None
    for attr in ("rope_theta",):
        theta = getattr(cfg, attr, None) if cfg is not None else None
    if theta is None and cfg is not None:
        for attr in ("rope_parameters", "rope_scaling"):
            rp = getattr(cfg, attr, None)
            if isinstance(rp, dict) and rp.get("rope_theta") is not None:
                theta = rp["rope_theta"]
    return int(heads), int(head_dim), float(theta if theta is not None else 10000.0)


def new_forward(self, hidden_states: torch.Tensor, attention_mask: Optional[torch.Tensor] = None,
                position_ids: Optional[torch.LongTensor] = None, past_key_value=None,
                output_attentions: bool = False, use_cache: bool = False, cache_position=None, **kwargs):
    """Same contract as the reference's ``new_forward`` (sdpa_transform.py:12-91): returns
    ``(attn_output [B,S,h], None, past_key_value)``."""
    if output_attentions:
        raise NotImplementedError("output_attentions is not supported (broken in the reference too: quirk q12)")
    if past_key_value is not None:
        raise NotImplementedError("past_key_value is not served at the operator seam")
    if attention_mask is None or attention_mask.dim() != 4:
        raise Exception("attention_mask parameter was unavailable or invalid")
    B, S, hidden = hidden_states.shape
    if attention_mask.shape[-1] != S:
        raise NotImplementedError("sequence-parallel chunks are not served at the operator seam")
    H, D, theta = _geometry(self)
    plan = _mask_plan(self, attention_mask)
    dev = hidden_states.device
    x = hidden_states.reshape(B * S, hidden)
    if x.dtype != torch.bfloat16:
        raise TypeError("videogpt_b200 attention computes in bf16; cast the model with .to(torch.bfloat16)")
    st = self.__dict__.setdefault("_vgpt_state", {})
    if st.get("shape") != (B, S, H, D, str(dev)):
        st.clear()
        inv_freq = (1.0 / (theta ** (torch.arange(0, D, 2, dtype=torch.int64).float() / D))).to(dev)
        st.update(shape=(B, S, H, D, str(dev)), inv_freq=inv_freq, table=None,
                  kv=torch.zeros(2, B * plan.pages, H, PAGE_TOKENS, D, device=dev, dtype=torch.bfloat16),
                  qkv=torch.empty(B * S, 3 * H * D, device=dev, dtype=torch.bfloat16),
                  attn=torch.empty(B * S, H * D, device=dev, dtype=torch.bfloat16))
    max_pos = int(position_ids.max()) + 1
    if st["table"] is None or st["table"].shape[0] < max_pos:
        st["table"] = ops.rope_table(st["inv_freq"], max_pos, D)
    qkv, attn, kv = st["qkv"], st["attn"], st["kv"]
    ops.gemm(x.contiguous(), self.qkv_proj.weight, out=qkv)
    ops.rope_kv_append(qkv, position_ids.reshape(-1).to(torch.int32), plan.row_slot, st["table"], kv[0], kv[1], H, D)
    ops.attention(qkv[:, :H * D], attn, kv[0], kv[1], plan.page_table, plan.seqs, S, plan.q_code, plan.k_code,
                  plan.k_tile_minmax, H, D, 1.0 / math.sqrt(D))
    out = ops.gemm(attn, self.o_proj.weight)
    return out.view(B, S, hidden), None, past_key_value


def replace_attention(model):
    """Install the B200 attention forward on every Phi-3 attention module of ``model`` (any module
    with ``qkv_proj`` and ``o_proj`` projections), as the reference's ``replace_attention`` does."""
    n = 0
    for module in model.modules():
        if hasattr(module, "qkv_proj") and hasattr(module, "o_proj"):
            module.forward = MethodType(new_forward, module)
            n += 1
    if n == 0:
        raise ValueError("no Phi-3 attention modules (qkv_proj / o_proj) found")
    return model
