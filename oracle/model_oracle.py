"""Oracle restatement of the reference's next-clip denoising forward (plain PyTorch).

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).  Runs on CPU in any dtype
(and, in ``-m gpu`` tests, on the GPU in bf16 as "the reference's own PyTorch
path on identical weights", which is the floating-point parity target named in
BASELINE.json).  Weights are a plain ``dict`` with the reference's state-dict
names.

The reference computes everything "as written": no KV cache (``LVM/scheduler.py:174``
passes ``past_key_values=None`` every step), the unconditional row left-padded to
the full length, a dense additive mask.  This file does the same on purpose.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Dict, List, Optional

import torch
import torch.nn.functional as F


@dataclass
class OracleConfig:
    hidden_size: int = 3072
    intermediate_size: int = 8192
    num_hidden_layers: int = 32
    num_attention_heads: int = 32
    rms_norm_eps: float = 1e-5
    rope_theta: float = 10000.0
    patch_size: int = 2
    in_channels: int = 4
    pos_embed_max_size: int = 192
    vocab_size: int = 32064

    @property
    def head_dim(self) -> int:
        return self.hidden_size // self.num_attention_heads


# --------------------------------------------------------------------------------------
# embedders (LVM/model.py)
# --------------------------------------------------------------------------------------

def sincos_pos_embed_table(embed_dim: int, grid: int, base_size: int = 64,
                           interpolation_scale: float = 1.0) -> torch.Tensor:
    """``get_2d_sincos_pos_embed`` (model.py:86-135) -> fp32 ``[grid*grid, embed_dim]``.

    Grid coordinate = index / (grid/base_size); meshgrid with w first (model.py:94-99);
    first half of the channels encodes ``grid[0]`` (the w coordinate), each half is
    ``[sin | cos]`` over ``omega_d = 1/10000^(d/(D/4))`` computed in float64."""
    import numpy as np
    gh = np.arange(grid, dtype=np.float32) / (grid / base_size) / interpolation_scale
    gw = np.arange(grid, dtype=np.float32) / (grid / base_size) / interpolation_scale
    mesh = np.stack(np.meshgrid(gw, gh), axis=0).reshape(2, 1, grid, grid)

    def one_dim(dim, pos):
        omega = np.arange(dim // 2, dtype=np.float64) / (dim / 2.0)
        omega = 1.0 / 10000 ** omega
        out = np.einsum("m,d->md", pos.reshape(-1), omega)
        return np.concatenate([np.sin(out), np.cos(out)], axis=1)

    emb = np.concatenate([one_dim(embed_dim // 2, mesh[0]), one_dim(embed_dim // 2, mesh[1])], axis=1)
    return torch.from_numpy(emb).float()


def cropped_pos_embed(pos_embed: torch.Tensor, max_size: int, height: int, width: int, patch: int):
    """``cropped_pos_embed`` (model.py:268-289): centre crop of the ``[1, max*max, h]`` buffer."""
    hh, ww = height // patch, width // patch
    if hh > max_size or ww > max_size:
        raise ValueError("latent larger than pos_embed_max_size")
    top, left = (max_size - hh) // 2, (max_size - ww) // 2
    pe = pos_embed.reshape(1, max_size, max_size, -1)[:, top:top + hh, left:left + ww, :]
    return pe.reshape(1, hh * ww, -1)


def patch_embed(latent, weight, bias, pos_embed, cfg: OracleConfig):
    """``PatchEmbedMR`` + pos-embed add (model.py:149-153, 300-306).  latent ``[1,C,H,W]``."""
    x = F.conv2d(latent, weight, bias, stride=cfg.patch_size)
    x = x.flatten(2).transpose(1, 2)
    return x + cropped_pos_embed(pos_embed, cfg.pos_embed_max_size, latent.shape[-2],
                                 latent.shape[-1], cfg.patch_size)


def timestep_embedding(t, dim: int = 256, max_period: float = 10000.0):
    """``TimestepEmbedder.timestep_embedding`` (model.py:39-58): ``[cos | sin]``, fp32."""
    half = dim // 2
    freqs = torch.exp(-math.log(max_period) * torch.arange(half, dtype=torch.float32) / half).to(t.device)
    args = t[:, None].float() * freqs[None]
    return torch.cat([torch.cos(args), torch.sin(args)], dim=-1)


def timestep_embedder(w, prefix: str, t, dtype):
    """``TimestepEmbedder.forward`` (model.py:60-63): Linear -> SiLU -> Linear."""
    x = timestep_embedding(t).to(dtype)
    x = F.linear(x, w[f"{prefix}.mlp.0.weight"], w[f"{prefix}.mlp.0.bias"])
    x = F.silu(x)
    return F.linear(x, w[f"{prefix}.mlp.2.weight"], w[f"{prefix}.mlp.2.bias"])


def final_layer(w, x, c):
    """``FinalLayer.forward`` (model.py:79-83): adaLN-modulated LayerNorm + Linear(h -> p*p*C)."""
    mod = F.linear(F.silu(c), w["final_layer.adaLN_modulation.1.weight"],
                   w["final_layer.adaLN_modulation.1.bias"])
    shift, scale = mod.chunk(2, dim=1)
    x = F.layer_norm(x, (x.shape[-1],), eps=1e-6)
    x = x * (1 + scale.unsqueeze(1)) + shift.unsqueeze(1)
    return F.linear(x, w["final_layer.linear.weight"], w["final_layer.linear.bias"])


def unpatchify(x, h: int, w: int, cfg: OracleConfig):
    """``unpatchify`` (model.py:255-265): feature order (p, q, c) -> ``[n, c, h, w]``."""
    p, c = cfg.patch_size, cfg.in_channels
    x = x.reshape(x.shape[0], h // p, w // p, p, p, c)
    x = torch.einsum("nhwpqc->nchpwq", x)
    return x.reshape(x.shape[0], c, h, w)


# --------------------------------------------------------------------------------------
# Phi-3 backbone (OmniGen/transformer.py + transformers==4.47.1 + sdpa_transform.py)
# --------------------------------------------------------------------------------------

def rms_norm(x, weight, eps):
    """``Phi3RMSNorm`` (transformers 4.47.1): fp32 statistics, cast back, then ``weight *``."""
    dt = x.dtype
    xf = x.to(torch.float32)
    xf = xf * torch.rsqrt(xf.pow(2).mean(-1, keepdim=True) + eps)
    return weight * xf.to(dt)


def rope_cos_sin(position_ids, head_dim: int, theta: float, dtype):
    """``Phi3RotaryEmbedding.forward`` (4.47.1; called at sdpa_transform.py:52): fp32
    ``inv_freq (x) position``, ``cat(freqs, freqs)``, cos/sin cast to the model dtype."""
    inv_freq = 1.0 / (theta ** (torch.arange(0, head_dim, 2, dtype=torch.int64,
                                             device=position_ids.device).float() / head_dim))
    freqs = (inv_freq[None, :, None].float().expand(position_ids.shape[0], -1, 1)
             @ position_ids[:, None, :].float()).transpose(1, 2)
    emb = torch.cat((freqs, freqs), dim=-1)
    return emb.cos().to(dtype), emb.sin().to(dtype)


def rotate_half(x):
    h = x.shape[-1] // 2
    return torch.cat((-x[..., h:], x[..., :h]), dim=-1)


def additive_mask(attention_mask, dtype):
    """``Phi3Transformer.forward`` mask conversion (OmniGen/transformer.py:128-151)."""
    if attention_mask is None or attention_mask.dim() != 3:
        raise Exception("attention_mask parameter was unavailable or invalid")
    min_dtype = torch.finfo(dtype).min
    m = -1 * (attention_mask + -1) * min_dtype
    return m.unsqueeze(1).to(dtype)


def attention(w, i: int, x, add_mask, cos, sin, cfg: OracleConfig):
    """``new_forward`` (sdpa_transform.py:37-91) at SP=1, no cache: fused qkv, half-split RoPE,
    SDPA with the dense additive mask (default scale 1/sqrt(d)), o_proj."""
    b, s, _ = x.shape
    nh, hd = cfg.num_attention_heads, cfg.head_dim
    qkv = F.linear(x, w[f"llm.layers.{i}.self_attn.qkv_proj.weight"])
    q, k, v = qkv.split(nh * hd, dim=-1)
    q = q.view(b, s, nh, hd).transpose(1, 2)
    k = k.view(b, s, nh, hd).transpose(1, 2)
    v = v.view(b, s, nh, hd).transpose(1, 2)
    c, sn = cos.unsqueeze(1), sin.unsqueeze(1)
    q = (q * c) + (rotate_half(q) * sn)
    k = (k * c) + (rotate_half(k) * sn)
    o = F.scaled_dot_product_attention(q.contiguous(), k.contiguous(), v.contiguous(),
                                       attn_mask=add_mask.to(q.dtype), dropout_p=0.0, is_causal=False)
    o = o.transpose(1, 2).contiguous().view(b, s, nh * hd)
    return F.linear(o, w[f"llm.layers.{i}.self_attn.o_proj.weight"])


def mlp(w, i: int, x):
    """``Phi3MLP.forward`` (4.47.1): ``down(up * silu(gate))`` with ``gate, up = chunk(gate_up)``."""
    gu = F.linear(x, w[f"llm.layers.{i}.mlp.gate_up_proj.weight"])
    gate, up = gu.chunk(2, dim=-1)
    return F.linear(up * F.silu(gate), w[f"llm.layers.{i}.mlp.down_proj.weight"])


def backbone(w, inputs_embeds, attention_mask, position_ids, cfg: OracleConfig,
             return_layer_outputs: bool = False):
    """``Phi3Transformer.forward`` (OmniGen/transformer.py:71-232) with the 4.47.1
    ``Phi3DecoderLayer`` (pre-norm attention + residual, pre-norm MLP + residual), final norm."""
    add_mask = additive_mask(attention_mask, inputs_embeds.dtype)
    cos, sin = rope_cos_sin(position_ids, cfg.head_dim, cfg.rope_theta, inputs_embeds.dtype)
    h = inputs_embeds
    per_layer = []
    for i in range(cfg.num_hidden_layers):
        r = h
        h = rms_norm(h, w[f"llm.layers.{i}.input_layernorm.weight"], cfg.rms_norm_eps)
        h = r + attention(w, i, h, add_mask, cos, sin, cfg)
        r = h
        h = rms_norm(h, w[f"llm.layers.{i}.post_attention_layernorm.weight"], cfg.rms_norm_eps)
        h = r + mlp(w, i, h)
        if return_layer_outputs:
            per_layer.append(h)
    h = rms_norm(h, w["llm.norm.weight"], cfg.rms_norm_eps)
    return (h, per_layer) if return_layer_outputs else h


# --------------------------------------------------------------------------------------
# LVM.frame_block_forward[_with_cfg] (LVM/model.py:399-566)
# --------------------------------------------------------------------------------------

def frame_block_forward(w, cfg: OracleConfig, x: List[torch.Tensor], timestep, input_ids,
                        input_img_latents, input_image_sizes, attention_mask, position_ids,
                        denoise_image_sizes, time_emb_inx, return_hidden: bool = False,
                        layer_outputs: Optional[list] = None):
    dtype = x[0].dtype
    shapes = [list(l.shape[-2:]) for l in x]
    xs = [patch_embed(l, w["x_embedder.proj.weight"], w["x_embedder.proj.bias"], w["pos_embed"], cfg)
          for l in x]                                                      # model.py:419
    time_token = timestep_embedder(w, "time_token", timestep, dtype)       # 420
    ctx = [patch_embed(l, w["input_x_embedder.proj.weight"], w["input_x_embedder.proj.bias"],
                       w["pos_embed"], cfg) for l in (input_img_latents or [])]   # 423
    emb = F.embedding(input_ids, w["llm.embed_tokens.weight"]).clone()     # 430-432
    n = 0
    for b in input_image_sizes.keys():                                     # 436-439
        for s, e in input_image_sizes[b]:
            emb[b, s:e] = ctx[n]; n += 1
    assert n == len(ctx)
    n = 0
    for b in time_emb_inx.keys():                                          # 443-446
        for tinx in time_emb_inx[b]:
            emb[b, tinx] = time_token[n]; n += 1
    assert n == time_token.shape[0]
    n = 0
    for b in denoise_image_sizes.keys():                                   # 450-453
        for s, e in denoise_image_sizes[b]:
            emb[b, s:e] = xs[n]; n += 1
    assert n == len(xs)
    if layer_outputs is not None:          # diagnostic: hidden state after every decoder layer
        hidden, per_layer = backbone(w, emb, attention_mask, position_ids, cfg, return_layer_outputs=True)
        layer_outputs.extend(per_layer)
    else:
        hidden = backbone(w, emb, attention_mask, position_ids, cfg)       # 465
    t_emb = timestep_embedder(w, "t_embedder", timestep, dtype)            # 480
    out, n = [], 0
    for b in denoise_image_sizes.keys():                                   # 481-486
        for s, e in denoise_image_sizes[b]:
            y = final_layer(w, hidden[b:b + 1, s:e], t_emb[n:n + 1])
            out.append(unpatchify(y, shapes[n][0], shapes[n][1], cfg))
            n += 1
    return (out, hidden) if return_hidden else out


def frame_block_forward_with_cfg(w, cfg: OracleConfig, x, timestep, *, input_ids, input_img_latents,
                                 input_image_sizes, attention_mask, position_ids, denoise_image_sizes,
                                 time_emb_inx, use_img_cfg, img_cfg_scale, prediction_type="v", **_):
    """``frame_block_forward_with_cfg`` (model.py:518-566): CFG inside the model only in v mode;
    returns the cond list duplicated (``cond + cond``)."""
    out = frame_block_forward(w, cfg, x, timestep, input_ids, input_img_latents, input_image_sizes,
                              attention_mask, position_ids, denoise_image_sizes, time_emb_inx)
    if use_img_cfg and prediction_type == "v":
        half = len(out) // 2
        cond, uncond = out[:half], out[half:]
        cond = [u + img_cfg_scale * (c - u) for c, u in zip(cond, uncond)]
        out = cond + cond
    return out


# --------------------------------------------------------------------------------------
# LVM.forward[_with_cfg] (LVM/model.py:330-397, 503-516): one output frame, batched tensor x
# --------------------------------------------------------------------------------------

def single_frame_forward(w, cfg: OracleConfig, x: torch.Tensor, timestep, input_ids, input_img_latents,
                         input_image_sizes, attention_mask, position_ids):
    dtype = x.dtype
    h_, w_ = x.shape[-2:]
    xe = patch_embed(x, w["x_embedder.proj.weight"], w["x_embedder.proj.bias"], w["pos_embed"], cfg)
    n_tok = xe.shape[1]
    time_token = timestep_embedder(w, "time_token", timestep, dtype).unsqueeze(1)
    ctx = [patch_embed(l, w["input_x_embedder.proj.weight"], w["input_x_embedder.proj.bias"],
                       w["pos_embed"], cfg) for l in (input_img_latents or [])]
    cond = F.embedding(input_ids, w["llm.embed_tokens.weight"]).clone()
    n = 0
    for b in input_image_sizes.keys():
        for s, e in input_image_sizes[b]:
            cond[b, s:e] = ctx[n]; n += 1
    emb = torch.cat([cond, time_token, xe], dim=1)
    hidden = backbone(w, emb, attention_mask, position_ids, cfg)
    t_emb = timestep_embedder(w, "t_embedder", timestep, dtype)
    y = final_layer(w, hidden[:, -n_tok:], t_emb)
    return unpatchify(y, h_, w_, cfg)


def single_frame_forward_with_cfg(w, cfg, x, timestep, *, input_ids, input_img_latents, input_image_sizes,
                                  attention_mask, position_ids, use_img_cfg, img_cfg_scale,
                                  prediction_type="v", **_):
    out = single_frame_forward(w, cfg, x, timestep, input_ids, input_img_latents, input_image_sizes,
                               attention_mask, position_ids)
    if use_img_cfg and prediction_type == "v":
        cond, uncond = torch.split(out, len(out) // 2, dim=0)
        cond = uncond + img_cfg_scale * (cond - uncond)
        out = torch.cat([cond, cond], dim=0)
    return out
