"""Tensor-level wrappers over the C ABI (include/vgpt_b200.h).

PyTorch is plumbing here: it owns device memory and the current stream; every wrapper
validates its tensors, passes raw pointers + sizes to ``libvgpt_b200.so`` on
``torch.cuda.current_stream()`` and raises on failure.  No wrapper computes anything on
the host or falls back to a torch implementation.
"""
from __future__ import annotations

import ctypes

import torch

from . import _lib

EPI_STORE, EPI_RESIDUAL, EPI_SWIGLU = 0, 1, 2
ROW_TOKEN, ROW_TIME, ROW_NOISY_PATCH, ROW_CONTEXT_PATCH = 0, 1, 2, 3
PAGE_TOKENS = 128
ATTN_KV_TILE = 64

BF16, I32, F32 = torch.bfloat16, torch.int32, torch.float32


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _p(t):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def _req(t: torch.Tensor, dtype, name: str, contiguous: bool = True):
    if not torch.is_tensor(t):
        raise TypeError(f"{name}: expected a tensor")
    if not t.is_cuda:
        raise RuntimeError(f"{name}: videogpt_b200 kernels need CUDA tensors (no CPU fallback), got {t.device}")
    if t.dtype != dtype:
        raise TypeError(f"{name}: expected {dtype}, got {t.dtype}")
    if contiguous and not t.is_contiguous():
        raise ValueError(f"{name}: tensor must be contiguous")
    return t


def gemm(a, w, out=None, residual=None, epilogue: int = EPI_STORE, block_n: int = 0, tail_mode: int = -1):
    """``out[M,N(/2)] = a[M,K] @ w[N,K]^T`` (+ epilogue).  ``a`` may be row-strided.  ``tail_mode``: -1 tuned default,
    1 plain 256-row tiles, 3 the ``M % 256 <= 32`` tail rows inside the k-loop of the last full tile row (same bits)."""
    _req(a, BF16, "a", contiguous=False)
    _req(w, BF16, "w")
    assert a.dim() == 2 and w.dim() == 2 and a.stride(1) == 1 and a.shape[1] == w.shape[1]
    M, K = a.shape
    N = w.shape[0]
    n_out = N // 2 if epilogue == EPI_SWIGLU else N
    if out is None:
        out = torch.empty(M, n_out, device=a.device, dtype=BF16)
    _req(out, BF16, "out", contiguous=False)
    assert out.shape == (M, n_out) and out.stride(1) == 1
    if residual is not None:
        _req(residual, BF16, "residual", contiguous=False)
        assert residual.shape == out.shape and residual.stride(0) == out.stride(0)
    _lib.call("vgpt_gemm_bf16", _p(a), _p(w), _p(out), _p(residual), M, N, K, a.stride(0),
              out.stride(0), epilogue, block_n, tail_mode, _stream())
    return out


def pack_gate_up(w):
    _req(w, BF16, "gate_up_proj.weight")
    two_i, k = w.shape
    out = torch.empty_like(w)
    _lib.call("vgpt_pack_gate_up", _p(w), _p(out), two_i // 2, k, _stream())
    return out


def rmsnorm(x, weight, eps: float, out=None):
    _req(x, BF16, "x"); _req(weight, BF16, "weight")
    rows, hidden = x.shape
    if out is None:
        out = torch.empty_like(x)
    _req(out, BF16, "out")
    _lib.call("vgpt_rmsnorm", _p(x), _p(weight), _p(out), rows, hidden, float(eps), _stream())
    return out


def rope_table(inv_freq, max_pos: int, head_dim: int):
    _req(inv_freq, F32, "inv_freq")
    assert inv_freq.numel() == head_dim // 2
    tab = torch.empty(max_pos, head_dim, device=inv_freq.device, dtype=BF16)
    _lib.call("vgpt_rope_table", _p(inv_freq), _p(tab), max_pos, head_dim, _stream())
    return tab


def rope_kv_append(qkv, row_pos, row_slot, table, k_pool, v_pool, heads: int, head_dim: int):
    _req(qkv, BF16, "qkv"); _req(row_pos, I32, "row_pos"); _req(row_slot, I32, "row_slot")
    _req(table, BF16, "table"); _req(k_pool, BF16, "k_pool"); _req(v_pool, BF16, "v_pool")
    rows = qkv.shape[0]
    assert qkv.shape[1] == 3 * heads * head_dim and row_pos.numel() >= rows and row_slot.numel() >= rows
    _lib.call("vgpt_rope_kv_append", _p(qkv), _p(row_pos), _p(row_slot), _p(table), _p(k_pool),
              _p(v_pool), rows, heads, head_dim, _stream())


def rope_kv_append_peers(qkv, row_pos, row_slot, table, k_pool_ptrs, v_pool_ptrs, n_pools: int, heads: int,
                         head_dim: int):
    """``k_pool_ptrs`` / ``v_pool_ptrs``: ctypes ``void*[n_pools]`` (``peer.SharedBuffer.ptr_array``)."""
    _req(qkv, BF16, "qkv"); _req(row_pos, I32, "row_pos"); _req(row_slot, I32, "row_slot"); _req(table, BF16, "table")
    rows = qkv.shape[0]
    assert qkv.shape[1] == 3 * heads * head_dim and row_pos.numel() >= rows and row_slot.numel() >= rows
    _lib.call("vgpt_rope_kv_append_peers", _p(qkv), _p(row_pos), _p(row_slot), _p(table), k_pool_ptrs,
              v_pool_ptrs, n_pools, rows, heads, head_dim, _stream())


ATTN_IMPL = "tcgen05"       # "mma_sync" selects the legacy cross-check kernel (tests only)


def attention(q, out, k_pool, v_pool, page_table, seqs, max_q_rows: int, q_code, k_code, k_tile_minmax,
              heads: int, head_dim: int, scale: float, impl: str = None):
    """q: [rows, >= H*D] (row-strided view allowed, e.g. the q part of qkv); out: [rows, H*D];
    pools: [pages, H, 128, D]."""
    _req(q, BF16, "q", contiguous=False); _req(out, BF16, "out", contiguous=False)
    _req(k_pool, BF16, "k_pool"); _req(v_pool, BF16, "v_pool")
    _req(page_table, I32, "page_table"); _req(seqs, I32, "seqs"); _req(q_code, I32, "q_code")
    _req(k_code, I32, "k_code"); _req(k_tile_minmax, I32, "k_tile_minmax")
    num_seqs, max_pages = page_table.shape
    assert seqs.shape == (num_seqs, 4) and k_code.shape == (num_seqs, max_pages * PAGE_TOKENS)
    max_k_tiles = k_tile_minmax.shape[1]
    assert k_tile_minmax.shape == (num_seqs, max_k_tiles, 2)
    assert k_pool.shape == v_pool.shape and k_pool.shape[1:] == (heads, PAGE_TOKENS, head_dim)
    name = "vgpt_attn_clip_causal" if (impl or ATTN_IMPL) == "tcgen05" else "vgpt_attn_clip_causal_mma_sync"
    _lib.call(name, _p(q), q.stride(0), q.shape[0], _p(out), out.stride(0), _p(k_pool), _p(v_pool),
              k_pool.shape[0], _p(page_table), max_pages, _p(seqs), num_seqs, max_q_rows, _p(q_code),
              _p(k_code), _p(k_tile_minmax), max_k_tiles, heads, head_dim, float(scale), _stream())
    return out


def embed_assemble(hidden, row_kind, row_a, row_b, embed_tokens, time_tokens, z, ctx, lat_h, lat_w,
                   w_noisy, b_noisy, w_ctx, b_ctx, pos_rows):
    _req(hidden, BF16, "hidden")
    rows, hs = hidden.shape
    for n, t in (("row_kind", row_kind), ("row_a", row_a), ("row_b", row_b)):
        _req(t, I32, n)
        assert t.numel() >= rows
    for n, t in (("embed_tokens", embed_tokens), ("w_noisy", w_noisy), ("b_noisy", b_noisy),
                 ("w_ctx", w_ctx), ("b_ctx", b_ctx), ("pos_rows", pos_rows)):
        _req(t, BF16, n)
    for n, t in (("time_tokens", time_tokens), ("z", z), ("ctx", ctx)):
        if t is not None:
            _req(t, BF16, n)
    _lib.call("vgpt_embed_assemble", _p(hidden), rows, hs, _p(row_kind), _p(row_a), _p(row_b),
              _p(embed_tokens), _p(time_tokens), _p(z), _p(ctx), 4, lat_h, lat_w, _p(w_noisy), _p(b_noisy),
              _p(w_ctx), _p(b_ctx), _p(pos_rows), _stream())
    return hidden


def timestep_sinusoid(t, freqs, out=None):
    _req(t, F32, "t"); _req(freqs, F32, "freqs")
    n, dim = t.numel(), 2 * freqs.numel()
    if out is None:
        out = torch.empty(n, dim, device=t.device, dtype=BF16)
    _lib.call("vgpt_timestep_sinusoid", _p(t), _p(freqs), _p(_req(out, BF16, "out")), n, dim, _stream())
    return out


def linear_small(x, w, bias, pre_silu: bool = False, post_silu: bool = False, out=None):
    _req(x, BF16, "x"); _req(w, BF16, "w")
    if bias is not None:
        _req(bias, BF16, "bias")
    n, k = x.shape
    N = w.shape[0]
    assert w.shape[1] == k
    if out is None:
        out = torch.empty(n, N, device=x.device, dtype=BF16)
    _req(out, BF16, "out")
    for r0 in range(0, n, 16):
        r1 = min(n, r0 + 16)
        _lib.call("vgpt_linear_small", _p(x[r0:r1]), _p(w), _p(bias), _p(out[r0:r1]), r1 - r0, N, k,
                  int(pre_silu), int(post_silu), _stream())
    return out


def final_layer(hidden, lat_row0, mod, w, bias, pred, norm_weight=None, rms_eps: float = 0.0, euler=None):
    """FinalLayer + unpatchify.  ``norm_weight``: ``hidden`` is the raw residual stream, apply the final RMSNorm first.
    ``euler = (z, scalars_dev, use_cfg, x1_mode, vel_out or None)``: apply the step's scheduler update to ``z`` in the
    same launch (``vgpt_cfg_euler`` arithmetic, device scalars ``[1 - sigma, d sigma, guidance]``)."""
    _req(hidden, BF16, "hidden"); _req(lat_row0, I32, "lat_row0"); _req(mod, BF16, "mod")
    _req(w, BF16, "w"); _req(bias, BF16, "bias"); _req(pred, BF16, "pred")
    n_lat, c, lat_h, lat_w = pred.shape
    assert mod.shape == (n_lat, 2 * hidden.shape[1]) and lat_row0.numel() >= n_lat
    if norm_weight is not None:
        _req(norm_weight, BF16, "norm_weight")
        assert norm_weight.numel() == hidden.shape[1]
    z = sc = vel = None
    use_cfg = x1 = 0
    if euler is not None:
        z, sc, use_cfg, x1, vel = euler
        _req(z, BF16, "z"); _req(sc, F32, "scalars_dev")
        assert z.shape == pred.shape and sc.numel() >= 3
        if vel is not None:
            _req(vel, BF16, "vel_out")
            assert vel.numel() >= (z.numel() // 2 if use_cfg else z.numel())
    _lib.call("vgpt_final_layer", _p(hidden), hidden.shape[1], _p(norm_weight), float(rms_eps), _p(lat_row0), _p(mod),
              _p(w), _p(bias), _p(pred), n_lat, c, lat_h, lat_w, _p(z), _p(vel), _p(sc), int(bool(use_cfg)), int(bool(x1)),
              _stream())
    return pred


def final_layer_rows(hidden, row_kind, row_a, row_b, mod, w, bias, pred_ptrs, n_preds: int, lat_h: int, lat_w: int,
                     norm_weight=None, rms_eps: float = 0.0):
    """Row-driven final layer; ``pred_ptrs``: ctypes ``void*[n_preds]`` of ``[n_lat,4,lat_h,lat_w]`` buffers."""
    _req(hidden, BF16, "hidden"); _req(mod, BF16, "mod"); _req(w, BF16, "w"); _req(bias, BF16, "bias")
    rows = hidden.shape[0]
    for n, t in (("row_kind", row_kind), ("row_a", row_a), ("row_b", row_b)):
        _req(t, I32, n)
        assert t.numel() >= rows
    if norm_weight is not None:
        _req(norm_weight, BF16, "norm_weight")
    _lib.call("vgpt_final_layer_rows", _p(hidden), rows, hidden.shape[1], _p(norm_weight), float(rms_eps), _p(row_kind),
              _p(row_a), _p(row_b), _p(mod), _p(w), _p(bias), pred_ptrs, n_preds, 4, lat_h, lat_w, _stream())


def cfg_euler(z, pred, use_cfg: bool, x1_mode: bool, one_minus_sigma: float = 1.0, dsigma: float = 0.0,
              guidance: float = 1.0, scalars_dev=None, vel_out=None):
    """In place on ``z`` ([n, C, h, w], cond latents first then uncond)."""
    _req(z, BF16, "z"); _req(pred, BF16, "pred")
    assert z.shape == pred.shape
    half = z.numel() // 2 if use_cfg else z.numel()
    if scalars_dev is not None:
        _req(scalars_dev, F32, "scalars_dev")
    if vel_out is not None:
        _req(vel_out, BF16, "vel_out")
        assert vel_out.numel() >= half
    _lib.call("vgpt_cfg_euler", _p(z), _p(pred), _p(vel_out), half, int(use_cfg), int(x1_mode),
              float(one_minus_sigma), float(dsigma), float(guidance), _p(scalars_dev), _stream())
    return z


def cfg_combine(pred, guidance: float):
    _req(pred, BF16, "pred")
    _lib.call("vgpt_cfg_combine", _p(pred), pred.numel() // 2, float(guidance), _stream())
    return pred


def mask_from_codes(q_code, k_code):
    _req(q_code, I32, "q_code"); _req(k_code, I32, "k_code")
    out = torch.empty(q_code.numel(), k_code.numel(), device=q_code.device, dtype=torch.uint8)
    _lib.call("vgpt_mask_from_codes", _p(q_code), _p(k_code), _p(out), q_code.numel(), k_code.numel(), _stream())
    return out


def umma_probe(a_img, b_img, a_desc_base: int, b_desc_base: int, idesc: int, k_steps: int,
               a_step_bytes: int, b_step_bytes: int, n_cols: int):
    """Test hook (tests/test_umma_layouts.py): raw smem images (uint8 CUDA tensors) -> fp32 [128, n_cols]."""
    assert a_img.is_cuda and b_img.is_cuda and a_img.dtype == torch.uint8 and b_img.dtype == torch.uint8
    out = torch.zeros(128, n_cols, device=a_img.device, dtype=F32)
    _lib.call_probe("vgpt_debug_umma_probe", _p(a_img), a_img.numel(), _p(b_img), b_img.numel(),
              ctypes.c_uint64(a_desc_base), ctypes.c_uint64(b_desc_base), ctypes.c_uint32(idesc), k_steps,
              ctypes.c_uint32(a_step_bytes), ctypes.c_uint32(b_step_bytes), _p(out), n_cols, _stream())
    return out


def umma_probe_ts(a_words, b_img, b_desc_base: int, idesc: int, k_steps: int, b_step_bytes: int, n_cols: int):
    """Test hook: A operand in tensor memory.  a_words: int32 CUDA tensor [128, cols] (packed bf16x2)."""
    assert a_words.is_cuda and a_words.dtype == torch.int32 and a_words.shape[0] == 128 and a_words.is_contiguous()
    out = torch.zeros(128, n_cols, device=a_words.device, dtype=F32)
    _lib.call_probe("vgpt_debug_umma_probe_ts", _p(a_words), a_words.shape[1], _p(b_img), b_img.numel(),
              ctypes.c_uint64(b_desc_base), ctypes.c_uint32(idesc), k_steps, ctypes.c_uint32(b_step_bytes), _p(out),
              n_cols, _stream())
    return out
