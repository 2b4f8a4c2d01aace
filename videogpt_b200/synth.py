"""Deterministic synthetic weights and latents (there is no network for checkpoints).

Shapes and names follow the reference's state dict; the
distributions follow ``LVM.initialize_weights`` (``LVM/model.py:213-244``: xavier-uniform
embedders, N(0, 0.02) timestep MLPs) and the HF Phi-3 init (N(0, 0.02)), EXCEPT that
``final_layer`` is drawn N(0, 0.02) instead of zero -- the reference zero-initialises it
(``model.py:241-244``), which would make every output exactly 0 and parity vacuous.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Dict

import numpy as np
import torch


@dataclass(frozen=True)
class BackboneDims:
    hidden_size: int = 3072
    intermediate_size: int = 8192
    num_hidden_layers: int = 32
    num_attention_heads: int = 32
    rms_norm_eps: float = 1e-5
    rope_theta: float = 10000.0
    vocab_size: int = 32064
    patch_size: int = 2
    in_channels: int = 4
    pos_embed_max_size: int = 192

    @property
    def head_dim(self) -> int:
        return self.hidden_size // self.num_attention_heads

    def phi3_kwargs(self) -> dict:
        """Keyword arguments for ``transformers.Phi3Config``."""
        return dict(hidden_size=self.hidden_size, intermediate_size=self.intermediate_size,
                    num_hidden_layers=self.num_hidden_layers,
                    num_attention_heads=self.num_attention_heads,
                    num_key_value_heads=self.num_attention_heads,
                    rms_norm_eps=self.rms_norm_eps, vocab_size=self.vocab_size)


FULL_SIZE = BackboneDims()
# BASELINE.json configs[0]: "2 layers / hidden 512" (FFN 1024, 8 heads x 64 chosen in SURVEY 8(d))
REDUCED = BackboneDims(hidden_size=512, intermediate_size=1024, num_hidden_layers=2,
                       num_attention_heads=8)


def sincos_pos_embed_rows(embed_dim: int, rows: np.ndarray, cols: np.ndarray,
                          grid: int = 192, base_size: int = 64,
                          interpolation_scale: float = 1.0) -> np.ndarray:
    """Rows of the reference's 2-D sincos table (``get_2d_sincos_pos_embed``,
    ``model.py:86-135``) for grid cells ``(rows[i], cols[i])``, fp64 -> fp32.

    Table entry (r, c): first half of the channels encodes the *column* coordinate
    (meshgrid "w goes first", model.py:96), second half the row coordinate; each half is
    ``[sin(p*omega) | cos(p*omega)]`` with ``omega_d = 10000^(-d/(D/4))`` in float64 and the
    coordinate ``p = index / (grid/base_size) / interpolation_scale`` in float32."""
    assert embed_dim % 4 == 0
    scale = np.float32(grid / base_size)
    pr = (rows.astype(np.float32) / scale / np.float32(interpolation_scale)).astype(np.float32)
    pc = (cols.astype(np.float32) / scale / np.float32(interpolation_scale)).astype(np.float32)
    quarter = embed_dim // 4
    omega = np.arange(quarter, dtype=np.float64) / (embed_dim / 4.0)
    omega = 1.0 / 10000 ** omega

    def half(p):
        out = np.einsum("m,d->md", p.reshape(-1), omega)
        return np.concatenate([np.sin(out), np.cos(out)], axis=1)

    return np.concatenate([half(pc), half(pr)], axis=1).astype(np.float32)


def sincos_pos_embed_table(embed_dim: int, grid: int = 192, base_size: int = 64,
                           interpolation_scale: float = 1.0) -> torch.Tensor:
    """Full persistent ``pos_embed`` buffer ``[1, grid*grid, embed_dim]`` fp32 (model.py:185-186)."""
    r, c = np.meshgrid(np.arange(grid), np.arange(grid), indexing="ij")
    tab = sincos_pos_embed_rows(embed_dim, r.reshape(-1), c.reshape(-1), grid, base_size,
                                interpolation_scale)
    return torch.from_numpy(tab).unsqueeze(0)


def cropped_pos_embed_rows(embed_dim: int, height: int, width: int, patch: int = 2,
                           grid: int = 192, base_size: int = 64,
                           interpolation_scale: float = 1.0) -> torch.Tensor:
    """The centre crop ``cropped_pos_embed`` (model.py:268-289) returns for a latent of
    ``height x width``, computed directly: fp32 ``[height/p * width/p, embed_dim]``."""
    hh, ww = height // patch, width // patch
    if hh > grid:
        raise ValueError(f"Height ({hh}) cannot be greater than `pos_embed_max_size`: {grid}.")
    if ww > grid:
        raise ValueError(f"Width ({ww}) cannot be greater than `pos_embed_max_size`: {grid}.")
    top, left = (grid - hh) // 2, (grid - ww) // 2
    r, c = np.meshgrid(np.arange(top, top + hh), np.arange(left, left + ww), indexing="ij")
    return torch.from_numpy(sincos_pos_embed_rows(embed_dim, r.reshape(-1), c.reshape(-1), grid,
                                                  base_size, interpolation_scale))


def state_dict_shapes(d: BackboneDims, with_pos_embed: bool = True) -> Dict[str, tuple]:
    h, i, p, c = d.hidden_size, d.intermediate_size, d.patch_size, d.in_channels
    s = {}
    if with_pos_embed:
        s["pos_embed"] = (1, d.pos_embed_max_size ** 2, h)
    for e in ("x_embedder", "input_x_embedder"):
        s[f"{e}.proj.weight"] = (h, c, p, p)
        s[f"{e}.proj.bias"] = (h,)
    for e in ("time_token", "t_embedder"):
        s[f"{e}.mlp.0.weight"] = (h, 256)
        s[f"{e}.mlp.0.bias"] = (h,)
        s[f"{e}.mlp.2.weight"] = (h, h)
        s[f"{e}.mlp.2.bias"] = (h,)
    s["final_layer.linear.weight"] = (p * p * c, h)
    s["final_layer.linear.bias"] = (p * p * c,)
    s["final_layer.adaLN_modulation.1.weight"] = (2 * h, h)
    s["final_layer.adaLN_modulation.1.bias"] = (2 * h,)
    s["llm.embed_tokens.weight"] = (d.vocab_size, h)
    for n in range(d.num_hidden_layers):
        s[f"llm.layers.{n}.self_attn.qkv_proj.weight"] = (3 * h, h)
        s[f"llm.layers.{n}.self_attn.o_proj.weight"] = (h, h)
        s[f"llm.layers.{n}.mlp.gate_up_proj.weight"] = (2 * i, h)
        s[f"llm.layers.{n}.mlp.down_proj.weight"] = (h, i)
        s[f"llm.layers.{n}.input_layernorm.weight"] = (h,)
        s[f"llm.layers.{n}.post_attention_layernorm.weight"] = (h,)
    s["llm.norm.weight"] = (h,)
    return s


def init_state_dict(d: BackboneDims, seed: int = 0, device="cpu", dtype=torch.float32,
                    with_pos_embed: bool = True, norm_jitter: float = 0.1) -> Dict[str, torch.Tensor]:
    """Random-init weights with the reference's names.  Same (seed, device type) -> same
    values.  RMSNorm weights are 1 + U(-j, j) and biases N(0, 0.02) rather than the
    reference's constant 1 / 0 so that a kernel that drops them cannot pass parity."""
    g = torch.Generator(device=device).manual_seed(seed)
    out = {}
    for name, shape in state_dict_shapes(d, with_pos_embed).items():
        if name == "pos_embed":
            t = sincos_pos_embed_table(d.hidden_size, d.pos_embed_max_size).to(device)
        elif name.endswith("layernorm.weight") or name == "llm.norm.weight":
            t = 1.0 + norm_jitter * (2 * torch.rand(shape, generator=g, device=device) - 1)
        elif name.endswith(".bias"):
            t = 0.02 * torch.randn(shape, generator=g, device=device)
        elif "embedder.proj.weight" in name:
            fan_in, fan_out = shape[1] * shape[2] * shape[3], shape[0]
            a = math.sqrt(6.0 / (fan_in + fan_out))
            t = a * (2 * torch.rand(shape, generator=g, device=device) - 1)
        else:
            t = 0.02 * torch.randn(shape, generator=g, device=device)
        out[name] = t.to(dtype)
    return out


def synthetic_latents(n_frames: int, height: int, width: int, seed: int = 42, device="cpu",
                      dtype=torch.float32, channels: int = 4):
    """``n_frames`` latents ``[1, C, height/8, width/8]`` ~ N(0,1) (SDXL-VAE latents x 0.13025
    are ~unit scale), drawn frame by frame like ``LVM/pipeline.py:478-481``."""
    g = torch.Generator(device=device).manual_seed(seed)
    return [torch.randn(1, channels, height // 8, width // 8, generator=g, device=device).to(dtype)
            for _ in range(n_frames)]


class SingleIdTagTokenizer:
    """Stand-in for the released tokenizer (not in the reference checkout, not downloadable):
    BOS (1) followed by one id per ``<img>`` / ``</img>`` / ``<|diffusion|>`` tag, which is what
    ``LVMProcessor`` assumes of the real one (``LVM/processor.py:138-142, 513-515``)."""
    TAGS = {"<img>": 32001, "</img>": 32002, "<|diffusion|>": 32003}
    eos_token_id = 2

    class _Out:
        def __init__(self, ids):
            self.input_ids = ids

    def __call__(self, text):
        import re
        ids = [1]
        for t in re.findall(r"<img>|</img>|<\|diffusion\|>", text):
            ids.append(self.TAGS[t])
        return self._Out(ids)
