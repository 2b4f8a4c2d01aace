"""Per-kernel parity on the GPU: every C-ABI entry point against the oracle restatement of the
reference op it replaces (oracle/model_oracle.py), on the same seeded inputs.

Tolerances: integer / index work bit-exact.  bf16 outputs are compared with the oracle
evaluated in fp32 on the bf16-rounded inputs; the bound is a few bf16 ulps of the output
scale (bf16 has 8 significand bits: one rounding is <= 2^-9 relative), written per test.
"""
import math

import numpy as np
import os

import pytest
import torch

from oracle import model_oracle as mo, processor_oracle as po, scheduler_oracle as so
from videogpt_b200 import synth

pytestmark = pytest.mark.gpu

DEV = "cuda"
BF = torch.bfloat16


@pytest.fixture(scope="module")
def ops():
    from videogpt_b200 import ops as _ops
    return _ops


def _rand(shape, seed, scale=1.0, dtype=BF):
    g = torch.Generator(device=DEV).manual_seed(seed)
    return (scale * torch.randn(shape, generator=g, device=DEV)).to(dtype)


def _rel(a, b):
    a, b = a.double(), b.double()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


# ---------------------------------------------------------------------------------------------
# GEMM (tcgen05): C = A W^T, three epilogues, M tails, both tile widths
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("M,N,K", [(128, 256, 64), (128, 256, 512), (16, 512, 3072), (2064, 1536, 512),
                                   (300, 3072, 1024), (1032, 9216, 3072)])
@pytest.mark.parametrize("block_n,tail_mode", [(256, 1), (192, 1), (128, 1), (0, -1)])
def test_gemm_store(ops, M, N, K, block_n, tail_mode):
    a, w = _rand((M, K), 1), _rand((N, K), 2, 0.05)
    c = ops.gemm(a, w, block_n=block_n, tail_mode=tail_mode)
    ref = a.float() @ w.float().t()
    # fp32 accumulation of exact bf16 products, one final rounding: <= 2^-8 of the row scale
    err = (c.float() - ref).abs().max().item()
    assert err <= 2 ** -7 * ref.abs().max().item(), err
    assert _rel(c, ref) < 4e-3


@pytest.mark.parametrize("block_n,tail_mode", [(256, 1), (192, 1), (128, 1)])
def test_gemm_identity_layout(ops, block_n, tail_mode):
    """W = I picks out columns of A exactly: catches any operand layout / swizzle / pair-split mix-up."""
    M, K = 520, 384
    a = _rand((M, K), 3)
    w = torch.eye(K, device=DEV, dtype=BF)
    c = ops.gemm(a, w, block_n=block_n, tail_mode=tail_mode)
    assert torch.equal(c, a)
    perm = torch.randperm(K, generator=torch.Generator().manual_seed(0)).to(DEV)
    c = ops.gemm(a, w[perm], block_n=block_n, tail_mode=tail_mode)
    assert torch.equal(c, a[:, perm])


@pytest.mark.parametrize("tail_mode,block_n", [(-1, 0), (1, 256), (1, 192)])
@pytest.mark.parametrize("M,N,K", [(2064, 512, 1024), (130, 3072, 8192)])
def test_gemm_residual_in_place(ops, M, N, K, tail_mode, block_n):
    a, w, r = _rand((M, K), 4), _rand((N, K), 5, 0.03), _rand((M, N), 6)
    want = (a.float() @ w.float().t()).to(BF).float() + r.float()     # bf16(o_proj) + residual, rounded
    out = r.clone()
    ops.gemm(a, w, out=out, residual=out, epilogue=ops.EPI_RESIDUAL, block_n=block_n, tail_mode=tail_mode)
    assert (out.float() - want).abs().max().item() <= 2 ** -6 * want.abs().max().item()
    assert _rel(out, want) < 4e-3


@pytest.mark.parametrize("tail_mode,block_n", [(-1, 0), (1, 256), (1, 192)])
@pytest.mark.parametrize("M,I,K", [(2064, 1024, 512), (257, 8192, 3072)])
def test_gemm_swiglu_matches_phi3_mlp(ops, M, I, K, tail_mode, block_n):
    """gate_up GEMM + SwiGLU epilogue on the packed weight == Phi3MLP's chunk / silu / mul."""
    x, wgu = _rand((M, K), 7), _rand((2 * I, K), 8, 0.03)
    packed = ops.pack_gate_up(wgu)
    # packing is a pure row permutation
    blk = torch.arange(2 * I, device=DEV).view(-1, 32)          # [gate x 16 | up x 16] per 32 packed rows
    src = torch.where(blk % 32 < 16, (blk // 32) * 16 + blk % 32, I + (blk // 32) * 16 + blk % 32 - 16)
    assert torch.equal(packed, wgu[src.view(-1)])
    h = ops.gemm(x, packed, epilogue=ops.EPI_SWIGLU, block_n=block_n, tail_mode=tail_mode)
    gu = (x.float() @ wgu.float().t()).to(BF)
    gate, up = gu.chunk(2, dim=-1)
    want = (up * torch.nn.functional.silu(gate)).float()
    assert h.shape == (M, I)
    assert _rel(h, want) < 6e-3


@pytest.mark.timeout(300)
@pytest.mark.parametrize("block_n", [0, 256, 192, 128])
@pytest.mark.parametrize("M", [2064, 1032, 280, 8208, 258, 272, 100, 16, 384, 2121, 544])
def test_gemm_tail_rows_in_the_k_loop_match_plain_tiles_bit_exact(ops, M, block_n):
    """M = q*256 + tail with q >= 1, tail <= 32: the tail rows are computed inside the k-loop of the last full
    tile row (special pieces, operands swapped; tail_mode=3 forces it wherever the shape allows, and it is the
    default of the auto path).  They -- and the rows of the narrower special pieces -- must get the same BITS
    as from plain 256-row tiles (tail_mode=1), for all three epilogues: a row's result may not depend on which
    kind of tile computed it (sequence-parallel shards and the unsharded run group rows differently).  Shapes
    the special path does not take (tail > 32, M < 256) fall back to plain tiles and trivially agree."""
    K, N = 512, 1024
    a, w, r = _rand((M, K), 21), _rand((N, K), 22, 0.05), _rand((M, N), 23)
    kw1, kw3 = dict(block_n=block_n, tail_mode=1), dict(block_n=block_n, tail_mode=3)
    want = ops.gemm(a, w, **kw1)
    assert torch.equal(ops.gemm(a, w, **kw3), want)
    assert torch.equal(ops.gemm(a, w), want)                      # the default path
    o1, o2 = r.clone(), r.clone()
    ops.gemm(a, w, out=o1, residual=o1, epilogue=ops.EPI_RESIDUAL, **kw1)
    ops.gemm(a, w, out=o2, residual=o2, epilogue=ops.EPI_RESIDUAL, **kw3)
    assert torch.equal(o1, o2)
    packed = ops.pack_gate_up(w)
    assert torch.equal(ops.gemm(a, packed, epilogue=ops.EPI_SWIGLU, **kw3), ops.gemm(a, packed, epilogue=ops.EPI_SWIGLU, **kw1))
    ref = a.float() @ w.float().t()
    assert _rel(ops.gemm(a, w, **kw3), ref) < 4e-3


@pytest.mark.timeout(300)
def test_gemm_tail_in_loop_full_size_projections(ops):
    """The four projection shapes of the model at M = 2064 (cfg2) and at a sequence-parallel shard (M = 1040):
    special pieces with the tail in the k-loop vs plain tiles, bit for bit (N = 9216 -> 41 pieces of 224 + one of 32;
    N = 16384 SwiGLU -> 85 of 192 + one of 64; N = 3072 -> 16 of 192)."""
    for M in (2064, 1040):
        for N, K, epi in ((9216, 3072, ops.EPI_STORE), (3072, 3072, ops.EPI_RESIDUAL), (16384, 3072, ops.EPI_SWIGLU),
                          (3072, 8192, ops.EPI_RESIDUAL)):
            a, w = _rand((M, K), 31), _rand((N, K), 32, 0.02)
            n_out = N // 2 if epi == ops.EPI_SWIGLU else N
            r = _rand((M, n_out), 33)
            o1, o2 = r.clone(), r.clone()
            res = dict(residual=o1) if epi == ops.EPI_RESIDUAL else {}
            ops.gemm(a, w, out=o1, epilogue=epi, tail_mode=1, **res)
            res = dict(residual=o2) if epi == ops.EPI_RESIDUAL else {}
            ops.gemm(a, w, out=o2, epilogue=epi, tail_mode=3, **res)
            assert torch.equal(o1, o2), (M, N, K, epi)


def test_gemm_rejects_bad_arguments(ops):
    from videogpt_b200._lib import VgptError
    with pytest.raises(VgptError):
        ops.gemm(_rand((8, 96), 1), _rand((64, 96), 2))          # K not a multiple of 64
    with pytest.raises(RuntimeError):
        ops.gemm(torch.zeros(8, 64, dtype=BF), torch.zeros(64, 64, dtype=BF))   # CPU tensors


# ---------------------------------------------------------------------------------------------
# elementwise kernels
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("rows,hidden", [(1, 512), (2064, 3072), (7, 4096)])
def test_rmsnorm(ops, rows, hidden):
    x, w = _rand((rows, hidden), 10, 3.0), (1 + 0.1 * _rand((hidden,), 11).float()).to(BF)
    y = ops.rmsnorm(x, w, 1e-5)
    want = mo.rms_norm(x, w, 1e-5)                   # bf16 oracle on the same device
    # identical rounding points; the variance reduction order may flip the last bit
    assert (y.float() - want.float()).abs().max().item() <= 2 ** -7 * want.float().abs().max().item()
    assert _rel(y, want) < 2e-3


@pytest.mark.parametrize("D", [64, 96])
def test_rope_table_and_kv_append(ops, D):
    H, rows, theta = 4, 300, 10000.0
    inv_freq = 1.0 / (theta ** (torch.arange(0, D, 2, dtype=torch.int64).float() / D))
    tab = ops.rope_table(inv_freq.to(DEV), 2100, D)
    pos = torch.arange(2100, device=DEV)[None]
    cos, sin = mo.rope_cos_sin(pos, D, theta, BF)
    assert torch.equal(tab[:, :D // 2], cos[0, :, :D // 2]) or \
        (tab[:, :D // 2].float() - cos[0, :, :D // 2].float()).abs().max().item() <= 2 ** -8
    assert (tab[:, D // 2:].float() - sin[0, :, :D // 2].float()).abs().max().item() <= 2 ** -8

    qkv = _rand((rows, 3 * H * D), 12)
    row_pos = torch.randint(0, 2100, (rows,), generator=torch.Generator().manual_seed(1)).to(DEV, torch.int32)
    n_pages = 4
    slots = torch.randperm(n_pages * 128, generator=torch.Generator().manual_seed(2))[:rows].to(DEV, torch.int32)
    slots[5] = -1
    k_pool = torch.zeros(n_pages, H, 128, D, device=DEV, dtype=BF)
    v_pool = torch.zeros_like(k_pool)
    q0, k0, v0 = [t.view(rows, H, D).clone() for t in qkv.split(H * D, dim=-1)]
    ops.rope_kv_append(qkv, row_pos, slots, tab, k_pool, v_pool, H, D)
    c = torch.cat([tab[:, :D // 2], tab[:, :D // 2]], -1)[row_pos.long()][:, None, :]
    s = torch.cat([tab[:, D // 2:], tab[:, D // 2:]], -1)[row_pos.long()][:, None, :]
    q_want = (q0 * c) + (mo.rotate_half(q0) * s)          # bf16 ops, as apply_rotary_pos_emb
    k_want = (k0 * c) + (mo.rotate_half(k0) * s)
    assert torch.equal(qkv[:, :H * D].view(rows, H, D), q_want)
    for r in range(rows):
        sl = int(slots[r])
        if sl < 0:
            continue
        assert torch.equal(k_pool[sl // 128, :, sl % 128], k_want[r]), r
        assert torch.equal(v_pool[sl // 128, :, sl % 128], v0[r]), r
    used = torch.zeros(n_pages * 128, dtype=torch.bool, device=DEV)
    used[slots[slots >= 0].long()] = True
    assert not k_pool.permute(0, 2, 1, 3).reshape(-1, H * D)[~used].any()      # nothing else written


def test_timestep_embedders_and_adaln(ops):
    d = synth.REDUCED
    sd = {k: v.to(DEV, BF) for k, v in synth.init_state_dict(d, seed=0, with_pos_embed=False).items()}
    t = torch.tensor([0.0, 0.02, 0.25, 0.5, 0.98, 1.0, 0.3333, 0.75], device=DEV)
    freqs = torch.exp(-math.log(10000.0) * torch.arange(128, dtype=torch.float32) / 128).to(DEV)
    sin = ops.timestep_sinusoid(t, freqs)
    want = mo.timestep_embedding(t).to(BF)
    assert (sin.float() - want.float()).abs().max().item() <= 2 ** -8
    for prefix in ("time_token", "t_embedder"):
        h1 = ops.linear_small(sin, sd[f"{prefix}.mlp.0.weight"], sd[f"{prefix}.mlp.0.bias"], post_silu=True)
        out = ops.linear_small(h1, sd[f"{prefix}.mlp.2.weight"], sd[f"{prefix}.mlp.2.bias"])
        ref = mo.timestep_embedder({k: v.float() for k, v in sd.items()}, prefix, t, torch.float32)
        assert _rel(out, ref) < 1e-2, prefix
    c = _rand((8, d.hidden_size), 20)
    mod = ops.linear_small(c, sd["final_layer.adaLN_modulation.1.weight"], sd["final_layer.adaLN_modulation.1.bias"],
                           pre_silu=True)
    ref = torch.nn.functional.linear(torch.nn.functional.silu(c.float()),
                                     sd["final_layer.adaLN_modulation.1.weight"].float(),
                                     sd["final_layer.adaLN_modulation.1.bias"].float())
    assert _rel(mod, ref) < 1e-2
    big = ops.linear_small(_rand((37, 256), 21), sd["t_embedder.mlp.0.weight"], None)     # > 16 rows: chunked
    assert big.shape == (37, d.hidden_size)


def test_embed_assemble_matches_patch_embed(ops):
    d = synth.REDUCED
    sd = {k: v.to(DEV) for k, v in synth.init_state_dict(d, seed=0).items()}
    sdb = {k: v.to(BF) for k, v in sd.items()}
    lat_h, lat_w = 8, 12
    z = _rand((3, 4, lat_h, lat_w), 30)
    ctx = _rand((2, 4, lat_h, lat_w), 31)
    time_tokens = _rand((3, d.hidden_size), 32)
    n_tok = (lat_h // 2) * (lat_w // 2)
    pos_rows = synth.cropped_pos_embed_rows(d.hidden_size, lat_h, lat_w).to(DEV, BF)
    cfg = mo.OracleConfig(hidden_size=d.hidden_size, intermediate_size=d.intermediate_size,
                          num_hidden_layers=d.num_hidden_layers, num_attention_heads=d.num_attention_heads)
    want_rows, kind, a, b = [], [], [], []
    for j in range(3):          # noisy latents
        e = mo.patch_embed(z[j:j + 1], sdb["x_embedder.proj.weight"], sdb["x_embedder.proj.bias"], sdb["pos_embed"], cfg)
        want_rows.append(e[0]); kind += [2] * n_tok; a += [j] * n_tok; b += list(range(n_tok))
    for j in range(2):          # context latents
        e = mo.patch_embed(ctx[j:j + 1], sdb["input_x_embedder.proj.weight"], sdb["input_x_embedder.proj.bias"],
                           sdb["pos_embed"], cfg)
        want_rows.append(e[0]); kind += [3] * n_tok; a += [j] * n_tok; b += list(range(n_tok))
    ids = [32001, 2, 0, 32003]
    want_rows.append(sdb["llm.embed_tokens.weight"][ids]); kind += [0] * 4; a += ids; b += [0] * 4
    want_rows.append(time_tokens[[2, 0]]); kind += [1, 1]; a += [2, 0]; b += [0, 0]
    want = torch.cat(want_rows, 0)
    hidden = torch.empty(want.shape[0], d.hidden_size, device=DEV, dtype=BF)
    i32 = lambda x: torch.tensor(x, dtype=torch.int32, device=DEV)
    ops.embed_assemble(hidden, i32(kind), i32(a), i32(b), sdb["llm.embed_tokens.weight"], time_tokens, z, ctx,
                       lat_h, lat_w, sdb["x_embedder.proj.weight"], sdb["x_embedder.proj.bias"],
                       sdb["input_x_embedder.proj.weight"], sdb["input_x_embedder.proj.bias"], pos_rows)
    tail = 6
    assert torch.equal(hidden[-tail:], want[-tail:])                 # pure row copies: exact
    # conv (+bias) in fp32 then two bf16 roundings on both sides; the K=16 sum order may differ
    assert (hidden[:-tail].float() - want[:-tail].float()).abs().max().item() <= 2 ** -6
    assert _rel(hidden[:-tail], want[:-tail]) < 3e-3


def test_final_layer_and_unpatchify(ops):
    d = synth.REDUCED
    sd = {k: v.to(DEV, BF) for k, v in synth.init_state_dict(d, seed=0, with_pos_embed=False).items()}
    cfg = mo.OracleConfig(hidden_size=d.hidden_size)
    lat_h, lat_w, n_lat = 8, 12, 3
    n_tok = (lat_h // 2) * (lat_w // 2)
    rows = 5 + n_lat * (n_tok + 2)
    hidden = _rand((rows, d.hidden_size), 40, 2.0)
    c = _rand((n_lat, d.hidden_size), 41)
    mod = torch.nn.functional.linear(torch.nn.functional.silu(c), sd["final_layer.adaLN_modulation.1.weight"],
                                     sd["final_layer.adaLN_modulation.1.bias"])
    row0 = [5 + j * (n_tok + 2) + 2 for j in range(n_lat)]
    pred = torch.zeros(n_lat, 4, lat_h, lat_w, device=DEV, dtype=BF)
    ops.final_layer(hidden, torch.tensor(row0, dtype=torch.int32, device=DEV), mod, sd["final_layer.linear.weight"],
                    sd["final_layer.linear.bias"], pred)
    for j in range(n_lat):
        x = hidden[row0[j]:row0[j] + n_tok][None]
        y = mo.final_layer(sd, x, c[j:j + 1])                       # bf16 oracle, same device
        want = mo.unpatchify(y, lat_h, lat_w, cfg)
        assert _rel(pred[j:j + 1], want) < 8e-3, j
        y32 = mo.final_layer({k: v.float() for k, v in sd.items()}, x.float(), c[j:j + 1].float())
        assert _rel(pred[j:j + 1], mo.unpatchify(y32, lat_h, lat_w, cfg)) < 2e-2, j


@pytest.mark.parametrize("mode", ["x1", "v"])
@pytest.mark.parametrize("use_cfg", [True, False])
def test_final_layer_fused_with_norm_and_scheduler_update_is_bit_exact(ops, mode, use_cfg):
    """One launch (final RMSNorm + FinalLayer + unpatchify + x1 -> v / CFG / Euler) == the three kernels it replaces at
    the end of every Euler step, bit for bit: prediction, updated latents and applied velocity."""
    d = synth.REDUCED
    sd = {k: v.to(DEV, BF) for k, v in synth.init_state_dict(d, seed=0, with_pos_embed=False).items()}
    lat_h, lat_w, n_lat = 8, 12, 4
    n_tok = (lat_h // 2) * (lat_w // 2)
    rows = 3 + n_lat * (n_tok + 2)
    hidden = _rand((rows, d.hidden_size), 60, 2.0)
    norm_w = (1.0 + 0.1 * _rand((d.hidden_size,), 61).float()).to(BF)
    mod = _rand((n_lat, 2 * d.hidden_size), 62, 0.5)
    row0 = torch.tensor([3 + j * (n_tok + 2) + 2 for j in range(n_lat)], dtype=torch.int32, device=DEV)
    w, b = sd["final_layer.linear.weight"], sd["final_layer.linear.bias"]
    z0 = _rand((n_lat, 4, lat_h, lat_w), 63)
    if use_cfg:
        z0[n_lat // 2:] = z0[:n_lat // 2]
    scal = torch.tensor([0.7, 0.02, 1.5], dtype=torch.float32, device=DEV)
    n_half = n_lat // 2 if use_cfg else n_lat
    # three kernels
    pred_a, z_a, vel_a = torch.zeros_like(z0), z0.clone(), torch.zeros_like(z0[:n_half])
    ops.final_layer(ops.rmsnorm(hidden, norm_w, 1e-5), row0, mod, w, b, pred_a)
    ops.cfg_euler(z_a, pred_a, use_cfg, mode == "x1", scalars_dev=scal, vel_out=vel_a)
    # norm fused only
    pred_b = torch.zeros_like(z0)
    ops.final_layer(hidden, row0, mod, w, b, pred_b, norm_weight=norm_w, rms_eps=1e-5)
    assert torch.equal(pred_b, pred_a)
    # everything in one launch
    pred_c, z_c, vel_c = torch.zeros_like(z0), z0.clone(), torch.zeros_like(z0[:n_half])
    ops.final_layer(hidden, row0, mod, w, b, pred_c, norm_weight=norm_w, rms_eps=1e-5, euler=(z_c, scal, use_cfg, mode == "x1", vel_c))
    torch.cuda.synchronize()
    assert torch.equal(pred_c, pred_a) and torch.equal(z_c, z_a) and torch.equal(vel_c, vel_a)


@pytest.mark.parametrize("mode", ["x1", "v"])
@pytest.mark.parametrize("use_cfg", [True, False])
def test_cfg_euler_matches_scheduler_oracle(ops, mode, use_cfg):
    n_cond, shape = 3, (1, 4, 8, 12)
    n = n_cond * (2 if use_cfg else 1)
    z = [_rand(shape, 50 + i) for i in range(n_cond)] * (2 if use_cfg else 1)
    pred = [_rand(shape, 60 + i) for i in range(n)]
    sigma = so.sigma_grid(50)
    i = 37
    g = 1.5

    def func(zz, t, prediction_type, **kw):          # a fixed "model output"; v-mode CFG as model.py:554-562
        out = [p.clone() for p in pred]
        if use_cfg and prediction_type == "v":
            h = len(out) // 2
            c = [u + g * (c - u) for c, u in zip(out[:h], out[h:])]
            out = c + c
        return out

    # one oracle step starting at sigma[i]
    s, s_next = sigma[i], sigma[i + 1]
    p = func(z, None, mode)
    if mode == "x1":
        p = [(a - b) / (1.0 - s) for a, b in zip(p, z)]
        if use_cfg:
            h = len(p) // 2
            c = [u + g * (c - u) for c, u in zip(p[:h], p[h:])]
            p = c + c
    want = [zz + (s_next - s) * pp for zz, pp in zip(z, p)]

    zt, pt = torch.cat(z, 0).clone(), torch.cat(pred, 0).clone()
    vel = torch.empty(n_cond, *shape[1:], device=DEV, dtype=BF)
    ops.cfg_euler(zt, pt, use_cfg, mode == "x1", float(1.0 - s), float(s_next - s), g, vel_out=vel)
    assert torch.equal(zt, torch.cat(want, 0))                       # same rounding points: bit-exact
    assert torch.equal(vel, torch.cat(p[:n_cond], 0))
    # scalars through device memory (CUDA-graph replay path)
    zt2 = torch.cat(z, 0).clone()
    sc = torch.tensor([float(1.0 - s), float(s_next - s), g], device=DEV)
    ops.cfg_euler(zt2, pt, use_cfg, mode == "x1", scalars_dev=sc)
    assert torch.equal(zt2, zt)


def test_cfg_combine(ops):
    c, u = _rand((3, 4, 8, 8), 70), _rand((3, 4, 8, 8), 71)
    pred = torch.cat([c, u], 0).clone()
    ops.cfg_combine(pred, 1.5)
    want = u + 1.5 * (c - u)
    assert torch.equal(pred[:3], want) and torch.equal(pred[3:], want)


# ---------------------------------------------------------------------------------------------
# mask: codes -> dense mask is bit-exact with the reference construction
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("case", [(4, 4, 256, 256, 1), (3, 4, 176, 320, 8), (5, 2, 64, 96, 8), (1, 1, 64, 64, 1)])
def test_mask_from_codes_bit_exact(ops, case):
    from videogpt_b200 import engine as eng
    n_ctx, n_gen, H, W, sp = case
    d = po.frame_block_inputs(n_ctx, n_gen, H, W, True, sp)
    specs, n_lat, n_c = eng.frame_block_specs(d["input_ids"], d["position_ids"], d["input_image_sizes"],
                                              d["denoise_image_sizes"], d["time_emb_inx"])
    assert n_lat == 2 * n_gen and n_c == n_ctx
    L = d["input_ids"].shape[1]
    for b, sp_ in enumerate(specs):
        pad = L - (sp_.n_prefix + sp_.n_active)
        qc = torch.from_numpy(np.concatenate([np.full(pad, eng.INT_MAX), sp_.codes]).astype(np.int32)).to(DEV)
        kc = torch.from_numpy(np.concatenate([np.full(pad, eng.INT_MAX - 1), sp_.codes]).astype(np.int32)).to(DEV)
        got = ops.mask_from_codes(qc, kc).bool().cpu()
        assert torch.equal(got, d["attention_mask"][b])
        assert torch.equal(torch.from_numpy(sp_.positions.astype(np.int64)), d["position_ids"][b, pad:])


# ---------------------------------------------------------------------------------------------
# attention over the paged cache vs dense-mask SDPA
# ---------------------------------------------------------------------------------------------
def _attention_case(ops, n_ctx, n_gen, H_px, W_px, heads, D, phase, seed, impl="tcgen05"):
    from videogpt_b200 import engine as eng
    d = po.frame_block_inputs(n_ctx, n_gen, H_px, W_px, True, 1)
    specs, n_lat, n_c = eng.frame_block_specs(d["input_ids"], d["position_ids"], d["input_image_sizes"],
                                              d["denoise_image_sizes"], d["time_emb_inx"])
    plan = eng.build_plan(specs, n_lat, n_c, H_px // 8, W_px // 8, DEV)
    L = d["input_ids"].shape[1]
    k_pool = _rand((plan.total_pages, heads, 128, D), seed + 1)       # garbage everywhere
    v_pool = _rand((plan.total_pages, heads, 128, D), seed + 2)
    ph = plan.prefix if phase == "prefix" else plan.step
    q = _rand((ph.rows, 3 * heads * D), seed + 3)                     # q lives in the fused qkv buffer
    out = torch.zeros(ph.rows, heads * D, device=DEV, dtype=BF)
    ops.attention(q[:, :heads * D], out, k_pool, v_pool, plan.page_table, ph.seqs, ph.max_q_rows, ph.q_code,
                  plan.k_code, plan.k_tile_minmax, heads, D, 1.0 / math.sqrt(D), impl=impl)
    # oracle: gather the logical K/V of each sequence and run dense-mask SDPA in fp32
    row0 = 0
    worst = 0.0
    for s, sp in enumerate(specs):
        T = sp.n_prefix + sp.n_active
        pad = L - T
        lo, hi = (0, sp.n_prefix) if phase == "prefix" else (sp.n_prefix, T)
        if hi == lo:
            continue
        logical = torch.arange(hi, device=DEV)
        pages = plan.page_table[s][logical // 128].long()
        K = k_pool[pages, :, logical % 128].float().permute(1, 0, 2)   # [H, kv, D]
        V = v_pool[pages, :, logical % 128].float().permute(1, 0, 2)
        Q = q[row0:row0 + hi - lo, :heads * D].float().view(hi - lo, heads, D).permute(1, 0, 2)
        mask = d["attention_mask"][s, pad + lo:pad + hi, pad:pad + hi].to(DEV)
        add = mo.additive_mask(mask[None], torch.float32)[0]
        want = torch.nn.functional.scaled_dot_product_attention(Q[None], K[None], V[None], attn_mask=add[None])[0]
        want = want.permute(1, 0, 2).reshape(hi - lo, heads * D)
        got = out[row0:row0 + hi - lo].float()
        worst = max(worst, _rel(got, want))
        # P and the output are rounded to bf16 once each: ~2^-8 relative to the value scale
        assert (got - want).abs().max().item() <= 3e-2 * want.abs().max().item()
        row0 += hi - lo
    assert worst < 1e-2, worst


IMPLS = ["tcgen05", "mma_sync"]


@pytest.mark.parametrize("impl", IMPLS)
@pytest.mark.parametrize("D", [64, 96, 128])
@pytest.mark.parametrize("phase", ["step", "prefix"])
def test_attention_small(ops, D, phase, impl):
    _attention_case(ops, 2, 2, 64, 96, 2, D, phase, 100, impl)


@pytest.mark.parametrize("impl", IMPLS)
@pytest.mark.parametrize("phase", ["step", "prefix"])
def test_attention_cfg2_geometry(ops, phase, impl):
    _attention_case(ops, 4, 4, 256, 256, 4, 96, phase, 200, impl)


@pytest.mark.parametrize("impl", IMPLS)
def test_attention_ragged_blocks(ops, impl):
    _attention_case(ops, 3, 5, 176, 320, 2, 96, "step", 300, impl)
    _attention_case(ops, 3, 5, 176, 320, 2, 96, "prefix", 301, impl)


def test_attention_long_context_tile_skipping(ops):
    """8 context clips (32 frames): most KV tiles of a context query tile are fully masked."""
    _attention_case(ops, 32, 4, 64, 64, 2, 96, "prefix", 400)
    _attention_case(ops, 32, 4, 64, 64, 2, 96, "step", 401)


@pytest.mark.timeout(300)
@pytest.mark.parametrize("phase", ["step", "prefix"])
def test_attention_cfg3_geometry(ops, phase):
    """BASELINE configs[2] at full frame size: 32 context frames of 258 tokens + 4 generated, 73 pages per
    sequence (the conditional CTA of a generated query pair walks 73 KV tiles, a context CTA up to 65)."""
    _attention_case(ops, 32, 4, 256, 256, 1, 96, phase, 450)
