"""CPU emulation of ``videogpt_b200.ops`` -- TEST INFRASTRUCTURE ONLY.

Every function has the signature of the C-ABI wrapper of the same name and does, with plain torch
on CPU, what the CUDA kernel of that name does ON THE SAME DATA LAYOUT: the paged K/V pools
``[page][H][128][D]`` addressed through ``row_slot`` / ``page_table``, the per-token codes instead
of a mask, the row kind / argument arrays of the assembly kernel, the ``lat_row0`` scatter of the
final layer, ``[cond..., uncond...]`` latents of the CFG/Euler kernel.  Monkeypatched over
``engine.ops`` / ``model.ops`` / ``scheduler.ops`` (fixture ``emu`` in ``tests/conftest.py``) it
lets the whole host side -- plan construction, prefix caching, page tables, latent numbering,
batching, rollout -- be checked NUMERICALLY against the oracle on a machine without a GPU, in fp32
(``engine.ACT_DTYPE`` patched), where a wrong slot, code or row shows up at 1e-1 and rounding at
1e-6.  It is not a fallback: nothing in ``videogpt_b200`` imports it, and the product's ``ops``
refuses CPU tensors.  The kernels themselves are checked against the oracle by the ``-m gpu`` tests.
"""
from __future__ import annotations


import torch
import torch.nn.functional as F

EPI_STORE, EPI_RESIDUAL, EPI_SWIGLU = 0, 1, 2
ROW_TOKEN, ROW_TIME, ROW_NOISY_PATCH, ROW_CONTEXT_PATCH = 0, 1, 2, 3
PAGE_TOKENS = 128
ATTN_KV_TILE = 64
INT_MAX = 2 ** 31 - 1

calls = []          # names of the emulated launches, in order (tests read and clear it)


def _log(name):
    calls.append(name)


def pack_gate_up(w):
    """The CUDA path block-interleaves gate/up rows for its epilogue; the emulation keeps the
    reference order and splits in ``gemm`` instead (same result)."""
    _log("pack_gate_up")
    return w.clone()


def gemm(a, w, out=None, residual=None, epilogue: int = EPI_STORE, block_n: int = 0, tail_mode: int = -1):
    _log("gemm")
    assert a.dim() == 2 and w.dim() == 2 and a.shape[1] == w.shape[1]
    y = a @ w.t()
    if epilogue == EPI_SWIGLU:
        gate, up = y.chunk(2, dim=-1)
        y = up * F.silu(gate)
    elif epilogue == EPI_RESIDUAL:
        assert residual is not None and residual.shape == y.shape
        y = y + residual
    else:
        assert residual is None
    if out is None:
        return y
    assert out.shape == y.shape
    out.copy_(y)
    return out


def rmsnorm(x, weight, eps: float, out=None):
    _log("rmsnorm")
    xf = x.float()
    y = weight * (xf * torch.rsqrt(xf.pow(2).mean(-1, keepdim=True) + eps)).to(x.dtype)
    if out is None:
        return y
    out.copy_(y)
    return out


def rope_table(inv_freq, max_pos: int, head_dim: int, dtype=None):
    """[max_pos, D] = cos(pos * inv_freq) | sin(pos * inv_freq), cast to the activation dtype."""
    _log("rope_table")
    from videogpt_b200 import engine
    ang = torch.arange(max_pos, dtype=torch.float32)[:, None] * inv_freq.float()[None, :]
    return torch.cat([ang.cos(), ang.sin()], dim=1).to(dtype or engine.ACT_DTYPE)


def rope_kv_append(qkv, row_pos, row_slot, table, k_pool, v_pool, heads: int, head_dim: int):
    _log("rope_kv_append")
    rows = qkv.shape[0]
    H, D, half = heads, head_dim, head_dim // 2
    assert qkv.shape[1] == 3 * H * D and row_pos.numel() >= rows and row_slot.numel() >= rows
    assert k_pool.shape == v_pool.shape and tuple(k_pool.shape[1:]) == (H, PAGE_TOKENS, D)
    if rows == 0:
        return
    pos = row_pos[:rows].long()
    assert int(pos.max()) < table.shape[0], "RoPE position beyond the table"
    cos, sin = table[pos, :half][:, None, :], table[pos, half:][:, None, :]

    def rot(x):                                   # [rows, H, D], half-split rotation
        lo, hi = x[..., :half], x[..., half:]
        return torch.cat([lo * cos - hi * sin, hi * cos + lo * sin], dim=-1)

    q = rot(qkv[:, :H * D].reshape(rows, H, D))
    k = rot(qkv[:, H * D:2 * H * D].reshape(rows, H, D))
    v = qkv[:, 2 * H * D:].reshape(rows, H, D)
    qkv[:, :H * D] = q.reshape(rows, H * D)
    slot = row_slot[:rows].long()
    keep = slot >= 0
    assert int(slot.max()) < k_pool.shape[0] * PAGE_TOKENS, "K/V slot beyond the pool"
    page, off = slot[keep] // PAGE_TOKENS, slot[keep] % PAGE_TOKENS
    k_pool[page, :, off, :] = k[keep]
    v_pool[page, :, off, :] = v[keep]


# ---- sequence parallel: the producing kernels store into every rank's buffers (peer.LocalPeerGroup) ----
_peer_bufs = []     # uint8 tensors of the virtual ranks' shared buffers (tests register them)


def register_peer_buffers(shared):
    """``shared``: a ``peer.SharedBuffer`` of a ``LocalPeerGroup`` on CPU -- remember every rank's bytes so
    that the raw pointers the engine hands to the ``*_peers`` wrappers can be turned back into tensors."""
    import ctypes
    n = shared.local.numel()
    for ptr in shared.ptrs:
        _peer_bufs[:] = [(b, t) for b, t in _peer_bufs if b != ptr]       # an address can be reused by a later test
        _peer_bufs.append((ptr, torch.frombuffer((ctypes.c_char * n).from_address(ptr), dtype=torch.uint8)))


def _from_ptr(ptr: int, dtype):
    for base, t in _peer_bufs:
        if base <= ptr < base + t.numel():
            return t[ptr - base:].view(dtype)
    raise AssertionError("pointer does not belong to a registered peer buffer")


def rope_kv_append_peers(qkv, row_pos, row_slot, table, k_pool_ptrs, v_pool_ptrs, n_pools: int, heads: int,
                         head_dim: int):
    per_page = heads * PAGE_TOKENS * head_dim
    q0 = qkv.clone()
    for g in range(n_pools):
        k = _from_ptr(int(k_pool_ptrs[g]), qkv.dtype)
        v = _from_ptr(int(v_pool_ptrs[g]), qkv.dtype)
        k = k[:k.numel() // per_page * per_page].view(-1, heads, PAGE_TOKENS, head_dim)
        v = v[:v.numel() // per_page * per_page].view(-1, heads, PAGE_TOKENS, head_dim)
        n = min(k.shape[0], v.shape[0])          # views run to the end of the buffer: trim to a common length
        k, v = k[:n], v[:n]
        qkv.copy_(q0)                            # q is rotated in place: once per call, not once per pool
        rope_kv_append(qkv, row_pos, row_slot, table, k, v, heads, head_dim)


def final_layer_rows(hidden, row_kind, row_a, row_b, mod, w, bias, pred_ptrs, n_preds: int, lat_h: int, lat_w: int,
                     norm_weight=None, rms_eps: float = 0.0):
    _log("final_layer_rows")
    if norm_weight is not None:
        hidden = rmsnorm(hidden, norm_weight, rms_eps)
        calls.pop()
    rows, hs = hidden.shape
    C, pw = 4, lat_w // 2
    preds = [_from_ptr(int(pred_ptrs[g]), hidden.dtype) for g in range(n_preds)]
    for r in range(rows):
        if int(row_kind[r]) != ROW_NOISY_PATCH:
            continue
        j, tkn = int(row_a[r]), int(row_b[r])
        y = _final_rows(hidden[r:r + 1], mod[j, :hs], mod[j, hs:], w, bias)[0].reshape(2, 2, C)     # (p, q, c)
        py, px = tkn // pw, tkn % pw
        for p_ in preds:
            view = p_[:(j + 1) * C * lat_h * lat_w].view(j + 1, C, lat_h, lat_w)
            view[j, :, 2 * py:2 * py + 2, 2 * px:2 * px + 2] = y.permute(2, 0, 1).to(view.dtype)


def attention(q, out, k_pool, v_pool, page_table, seqs, max_q_rows: int, q_code, k_code, k_tile_minmax,
              heads: int, head_dim: int, scale: float, impl: str = None):
    _log("attention")
    H, D = heads, head_dim
    num_seqs, max_pages = page_table.shape
    assert seqs.shape == (num_seqs, 4) and k_code.shape == (num_seqs, max_pages * PAGE_TOKENS)
    assert k_tile_minmax.shape[0] == num_seqs and k_tile_minmax.shape[2] == 2
    assert int(seqs[:, 1].max()) <= max_q_rows
    for s in range(num_seqs):
        row0, n, kv_len, _ = (int(x) for x in seqs[s])
        if n == 0:
            continue
        assert kv_len <= max_pages * PAGE_TOKENS
        logical = torch.arange(kv_len)
        pages = page_table[s, logical // PAGE_TOKENS].long()
        off = logical % PAGE_TOKENS
        k = k_pool[pages, :, off, :].float()            # [kv_len, H, D]
        v = v_pool[pages, :, off, :].float()
        kc = k_code[s, :kv_len].long()
        # the tile classification the kernel skips / fast-paths by must bound these codes: a tile is
        # skipped when no query reaches its minimum, and taken without the predicate when every
        # query reaches its maximum
        for t in range((kv_len + ATTN_KV_TILE - 1) // ATTN_KV_TILE):
            seg = k_code[s, t * ATTN_KV_TILE:min(kv_len, (t + 1) * ATTN_KV_TILE)]
            assert int(k_tile_minmax[s, t, 0]) <= int(seg.min()) and int(k_tile_minmax[s, t, 1]) >= int(seg.max()), \
                f"k_tile_minmax[{s},{t}] does not bound k_code"
        qq = q[row0:row0 + n, :H * D].reshape(n, H, D).float()
        allowed = q_code[row0:row0 + n].long()[:, None] >= kc[None, :]
        assert bool(allowed.any(dim=1).all()), "a query row sees no key"
        sc = torch.einsum("qhd,khd->hqk", qq, k) * scale
        sc = sc.masked_fill(~allowed[None], float("-inf"))
        p = torch.softmax(sc, dim=-1)
        o = torch.einsum("hqk,khd->qhd", p, v)
        out[row0:row0 + n, :H * D] = o.reshape(n, H * D).to(out.dtype)
    return out


def embed_assemble(hidden, row_kind, row_a, row_b, embed_tokens, time_tokens, z, ctx, lat_h, lat_w,
                   w_noisy, b_noisy, w_ctx, b_ctx, pos_rows):
    _log("embed_assemble")
    rows, hs = hidden.shape
    kind, a, b = row_kind[:rows].long(), row_a[:rows].long(), row_b[:rows].long()
    pw = lat_w // 2
    for r in range(rows):
        kd = int(kind[r])
        if kd == ROW_TOKEN:
            hidden[r] = embed_tokens[a[r]]
        elif kd == ROW_TIME:
            hidden[r] = time_tokens[a[r]]
        else:
            lat = (z if kd == ROW_NOISY_PATCH else ctx)[a[r]]            # [C, lat_h, lat_w]
            w, bias = (w_noisy, b_noisy) if kd == ROW_NOISY_PATCH else (w_ctx, b_ctx)
            py, px = int(b[r]) // pw, int(b[r]) % pw
            patch = lat[:, 2 * py:2 * py + 2, 2 * px:2 * px + 2].reshape(-1)   # (c, ph, pw) order
            conv = (w.reshape(hs, -1) @ patch + bias).to(hidden.dtype)
            hidden[r] = conv + pos_rows[b[r]]
    return hidden


def timestep_sinusoid(t, freqs, out=None):
    _log("timestep_sinusoid")
    args = t.float()[:, None] * freqs.float()[None]
    y = torch.cat([args.cos(), args.sin()], dim=-1)
    if out is None:
        return y
    out.copy_(y.to(out.dtype))
    return out


def linear_small(x, w, bias, pre_silu: bool = False, post_silu: bool = False, out=None):
    _log("linear_small")
    y = F.linear(F.silu(x) if pre_silu else x, w, bias)
    if post_silu:
        y = F.silu(y)
    if out is None:
        return y
    out.copy_(y)
    return out


def _final_rows(x, shift, scale, w, bias):
    y = F.layer_norm(x, (x.shape[-1],), eps=1e-6)
    y = y * (1 + scale) + shift
    return F.linear(y, w, bias)                      # [tokens, 16], feature order (p, q, c)


def final_layer(hidden, lat_row0, mod, w, bias, pred, norm_weight=None, rms_eps: float = 0.0, euler=None):
    _log("final_layer")
    if norm_weight is not None:
        hidden = rmsnorm(hidden, norm_weight, rms_eps)
        calls.pop()
    n_lat, C, lat_h, lat_w = pred.shape
    hs = hidden.shape[1]
    tokens = (lat_h // 2) * (lat_w // 2)
    assert mod.shape == (n_lat, 2 * hs)
    for j in range(n_lat):
        r0 = int(lat_row0[j])
        assert 0 <= r0 and r0 + tokens <= hidden.shape[0], "latent rows outside the hidden matrix"
        y = _final_rows(hidden[r0:r0 + tokens], mod[j, :hs], mod[j, hs:], w, bias)
        y = y.reshape(lat_h // 2, lat_w // 2, 2, 2, C)
        pred[j] = torch.einsum("hwpqc->chpwq", y).reshape(C, lat_h, lat_w).to(pred.dtype)
    if euler is not None:
        z, sc, use_cfg, x1, vel = euler
        cfg_euler(z, pred, bool(use_cfg), bool(x1), scalars_dev=sc, vel_out=vel)
        calls.pop()
    return pred


def cfg_euler(z, pred, use_cfg: bool, x1_mode: bool, one_minus_sigma: float = 1.0, dsigma: float = 0.0,
              guidance: float = 1.0, scalars_dev=None, vel_out=None):
    _log("cfg_euler")
    assert z.shape == pred.shape
    if scalars_dev is not None:
        one_minus_sigma, dsigma, guidance = (float(x) for x in scalars_dev[:3])
    n = z.shape[0]
    half = n // 2 if use_cfg else n
    c = pred[:half]
    if x1_mode:
        c = (c - z[:half]) * (1.0 / one_minus_sigma)
    if use_cfg:
        u = pred[half:]
        if x1_mode:
            u = (u - z[half:]) * (1.0 / one_minus_sigma)
        c = u + guidance * (c - u)
    if vel_out is not None:
        vel_out.copy_(c.reshape(vel_out.shape))
    z[:half] += dsigma * c
    if use_cfg:
        z[half:] += dsigma * c
    return z


def cfg_combine(pred, guidance: float):
    _log("cfg_combine")
    half = pred.shape[0] // 2
    c = pred[half:] + guidance * (pred[:half] - pred[half:])
    pred[:half] = c
    pred[half:] = c
    return pred


def mask_from_codes(q_code, k_code):
    _log("mask_from_codes")
    return (q_code[:, None] >= k_code[None, :]).to(torch.uint8)
