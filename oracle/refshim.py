"""Make the UNMODIFIED reference importable on CPU in the build container.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).  Used by
``tests/golden/make_golden.py`` and ``tests/test_oracle_vs_reference.py``; it is
a no-op on the GPU box, where ``/root/reference`` does not exist.

The reference (pure Python) pins ``transformers==4.47.1`` + diffusers/timm/peft/
deepspeed (``env_nv.sh:15-21``); this image has transformers 5.5 and none of
the others.  None of the missing packages is on the arithmetic path at
sequence-parallel size 1, so they are stubbed; the three 4.47.1 call
conventions the reference relies on are restored on top of the installed Phi-3
blocks (same arithmetic):

* ``Phi3DecoderLayer.forward`` returning a tuple and taking ``position_ids``
  (called at ``OmniGen/transformer.py:196-204``),
* ``apply_rotary_pos_emb(q, k, cos, sin, position_ids)`` and a per-attention
  ``rotary_emb(x, position_ids, seq_len=)`` (``LVM/transform/sdpa_transform.py:52-53``),
* ``dist_attn`` = the Ulysses wrapper, whose all-to-alls are the identity at
  P=1 (``sdpa_transform.py:94-159``).
"""
from __future__ import annotations

import importlib.machinery
import logging
import os
import re
import sys
import types
from types import MethodType

REFERENCE_ROOT = os.environ.get("VGPT_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "LVM"))


def _mod(name, **attrs):
    m = types.ModuleType(name)
    m.__spec__ = importlib.machinery.ModuleSpec(name, None)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


class _Dummy:
    def __init__(self, *a, **k):
        pass


_installed = False


def install():
    """Install the stubs and put the reference on sys.path (idempotent)."""
    global _installed
    if _installed:
        return
    if not reference_available():
        raise RuntimeError(f"reference checkout not found at {REFERENCE_ROOT}")

    class _Log:
        @staticmethod
        def get_logger(n):
            return logging.getLogger(n)

    for name in ("diffusers", "timm", "peft", "deepspeed"):
        if name in sys.modules:
            continue
    _mod("diffusers")
    _mod("diffusers.loaders", PeftAdapterMixin=type("PeftAdapterMixin", (object,), {}))
    _mod("diffusers.models", AutoencoderKL=_Dummy)
    _mod("diffusers.utils", USE_PEFT_BACKEND=False, is_torch_xla_available=lambda: False,
         logging=_Log, replace_example_docstring=lambda s: (lambda f: f),
         scale_lora_layers=None, unscale_lora_layers=None)
    _mod("diffusers.optimization", get_scheduler=None)
    _mod("timm")
    _mod("timm.models")
    _mod("timm.models.vision_transformer", PatchEmbed=_Dummy, Attention=_Dummy, Mlp=_Dummy)
    _mod("peft", LoraConfig=_Dummy, PeftModel=_Dummy)
    _mod("deepspeed", init_distributed=lambda *a, **k: None)
    _mod("deepspeed.sequence")
    _mod("deepspeed.sequence.layer", DistributedAttention=_Dummy, _SeqAllToAll=_Dummy)

    import transformers.cache_utils as cu
    if not hasattr(cu, "OffloadedCache"):
        cu.OffloadedCache = cu.DynamicCache
    import transformers.models.phi3.modeling_phi3 as mp
    if not hasattr(mp, "Phi3SdpaAttention"):
        mp.Phi3SdpaAttention = mp.Phi3Attention

    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    _installed = True


class FakeTokenizer:
    """One id per tag after BOS, as ``LVMProcessor`` assumes
    (``LVM/processor.py:138-142, 513-515``).  The released tokenizer is not in the
    reference checkout and cannot be downloaded."""

    TAGS = {"<img>": 32001, "</img>": 32002, "<|diffusion|>": 32003}
    eos_token_id = 2

    class _Out:
        def __init__(self, ids):
            self.input_ids = ids

    def __call__(self, text):
        ids = [1]
        for t in re.findall(r"<img>|</img>|<\|diffusion\|>", text):
            ids.append(self.TAGS[t])
        return self._Out(ids)


def _layer_forward_4471(self, hidden_states, attention_mask=None, position_ids=None,
                        past_key_value=None, output_attentions=False, use_cache=False,
                        cache_position=None, **kw):
    """transformers 4.47.1 ``Phi3DecoderLayer.forward`` call shape (tuple return)."""
    residual = hidden_states
    hidden_states = self.input_layernorm(hidden_states)
    attn_out, w, present = self.self_attn(
        hidden_states=hidden_states, attention_mask=attention_mask, position_ids=position_ids,
        past_key_value=past_key_value, output_attentions=output_attentions,
        use_cache=use_cache, cache_position=cache_position)
    hidden_states = residual + self.resid_attn_dropout(attn_out)
    residual = hidden_states
    hidden_states = self.post_attention_layernorm(hidden_states)
    hidden_states = self.mlp(hidden_states)
    hidden_states = residual + self.resid_mlp_dropout(hidden_states)
    out = (hidden_states,)
    if output_attentions:
        out += (w,)
    if use_cache:
        out += (present,)
    return out


def patch_llm(llm):
    """Wire the reference's own ``new_forward`` into every attention module of
    ``llm`` (an ``OmniGen.transformer.Phi3Transformer``) at SP=1."""
    install()
    import torch.nn.functional as F
    import transformers.models.phi3.modeling_phi3 as mp
    import LVM.transform.sdpa_transform as st

    new_rope = mp.apply_rotary_pos_emb
    st.apply_rotary_pos_emb = (
        lambda q, k, cos, sin, position_ids=None, unsqueeze_dim=1: new_rope(q, k, cos, sin, unsqueeze_dim))

    def local_dist_attn(q, k, v, batch_dim_idx, **kw):
        o = F.scaled_dot_product_attention(q.transpose(1, 2).contiguous(), k.transpose(1, 2).contiguous(),
                                           v.transpose(1, 2).contiguous(), **kw)
        return o.transpose(1, 2).contiguous()

    for layer in llm.layers:
        a = layer.self_attn
        a.num_heads = llm.config.num_attention_heads
        a.hidden_size = llm.config.hidden_size
        a.rotary_emb = (lambda x, position_ids, seq_len=None, _r=llm.rotary_emb: _r(x, position_ids))
        a.forward = MethodType(st.new_forward, a)
        a.dist_attn = local_dist_attn
        layer.forward = MethodType(_layer_forward_4471, layer)
    return llm


def build_reference_model(cfg_kwargs: dict, state_dict=None, dtype=None):
    """Construct the reference's own ``LVM`` on CPU and (optionally) load weights."""
    install()
    import torch
    from transformers import Phi3Config
    from LVM.acceleration.parallel_states import hccl_info
    from LVM.model import LVM

    hccl_info.world_size = 1
    hccl_info.rank = 0
    cfg = Phi3Config(**cfg_kwargs, use_cache=False)
    model = LVM(cfg).eval()
    patch_llm(model.llm)
    if state_dict is not None:
        missing, unexpected = model.load_state_dict(state_dict, strict=False)
        assert not unexpected, unexpected
        assert all("rotary" in m or "inv_freq" in m for m in missing), missing
    if dtype is not None:
        model.to(dtype)
    return model


def build_reference_processor(sequence_parallel_size: int = 1, max_image_size: int = 1024):
    install()
    from LVM.processor import LVMProcessor
    return LVMProcessor(FakeTokenizer(), max_image_size=max_image_size,
                        sequence_parallel_size=sequence_parallel_size)
