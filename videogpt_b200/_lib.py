"""ctypes binding of libvgpt_b200.so (C ABI: include/vgpt_b200.h).

There is NO fallback: if the library is missing or a call fails, an exception is raised.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import c_char_p, c_float, c_int, c_uint32, c_uint64, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libvgpt_b200.so")

P, I, F = c_void_p, c_int, c_float

# name -> argument types; every symbol declared in include/vgpt_b200.h
SIGNATURES = {
    "vgpt_abi_version": [],
    "vgpt_last_error": [],
    "vgpt_gemm_bf16": [P, P, P, P, I, I, I, I, I, I, I, I, P],
    "vgpt_pack_gate_up": [P, P, I, I, P],
    "vgpt_rmsnorm": [P, P, P, I, I, F, P],
    "vgpt_rope_table": [P, P, I, I, P],
    "vgpt_rope_kv_append": [P, P, P, P, P, P, I, I, I, P],
    "vgpt_rope_kv_append_peers": [P, P, P, P, P, P, I, I, I, I, P],
    "vgpt_final_layer_rows": [P, I, I, P, F, P, P, P, P, P, P, P, I, I, I, I, P],
    "vgpt_peer_alloc": [P, c_uint64],
    "vgpt_peer_free": [P],
    "vgpt_peer_export": [P, P],
    "vgpt_peer_import": [P, P],
    "vgpt_peer_close": [P],
    "vgpt_peer_barrier": [P, I, I, P, P],
    "vgpt_attn_clip_causal": [P, I, I, P, I, P, P, I, P, I, P, I, I, P, P, P, I, I, I, F, P],
    "vgpt_attn_clip_causal_mma_sync": [P, I, I, P, I, P, P, I, P, I, P, I, I, P, P, P, I, I, I, F, P],
    "vgpt_embed_assemble": [P, I, I, P, P, P, P, P, P, P, I, I, I, P, P, P, P, P, P],
    "vgpt_timestep_sinusoid": [P, P, P, I, I, P],
    "vgpt_linear_small": [P, P, P, P, I, I, I, I, I, P],
    "vgpt_final_layer": [P, I, P, F, P, P, P, P, P, I, I, I, I, P, P, P, I, I, P],
    "vgpt_cfg_euler": [P, P, P, I, I, I, F, F, F, P, P],
    "vgpt_cfg_combine": [P, I, F, P],
    "vgpt_mask_from_codes": [P, P, P, I, I, P],
    "vgpt_debug_attn_trace": [P, I, P, P],
}

# libvgpt_b200_probe.so (include/vgpt_b200_probe.h): descriptor / issue-rate probes for tests and tools, kept out of
# the product library
PROBE_LIB_PATH = os.path.join(_HERE, "libvgpt_b200_probe.so")
PROBE_SIGNATURES = {
    "vgpt_probe_last_error": [],
    "vgpt_debug_umma_rate": [I, I, I, I, I, I, P, P],
    "vgpt_debug_umma_probe_ts": [P, I, P, I, c_uint64, c_uint32, I, c_uint32, P, I, P],
    "vgpt_debug_umma_probe": [P, I, P, I, c_uint64, c_uint64, c_uint32, I, c_uint32, c_uint32, P, I, P],
}

_lib = None
_probe = None


class VgptError(RuntimeError):
    pass


def load(path: str = LIB_PATH) -> ctypes.CDLL:
    """Load the shared library and bind every exported symbol (raises if anything is missing)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(path):
        raise ImportError(
            f"{path} not found: the CUDA library has not been built. Run "
            "`python -m videogpt_b200.build` (needs nvcc, sm_100a). There is no CPU fallback.")
    lib = ctypes.CDLL(path)
    for name, argtypes in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError if the symbol is not exported
        fn.argtypes = argtypes
        fn.restype = c_char_p if name == "vgpt_last_error" else c_int
    _lib = lib
    return lib


def call(name: str, *args) -> None:
    """Call an int-returning entry point; raise VgptError on a non-zero return."""
    lib = load()
    rc = getattr(lib, name)(*args)
    if rc != 0:
        msg = lib.vgpt_last_error()
        raise VgptError(f"{name} failed (rc={rc}): {msg.decode() if msg else '?'}")


def load_probe(path: str = PROBE_LIB_PATH) -> ctypes.CDLL:
    global _probe
    if _probe is not None:
        return _probe
    if not os.path.exists(path):
        raise ImportError(f"{path} not found: run `python -m videogpt_b200.build`")
    lib = ctypes.CDLL(path)
    for name, argtypes in PROBE_SIGNATURES.items():
        fn = getattr(lib, name)
        fn.argtypes = argtypes
        fn.restype = c_char_p if name == "vgpt_probe_last_error" else c_int
    _probe = lib
    return lib


def call_probe(name: str, *args) -> None:
    lib = load_probe()
    rc = getattr(lib, name)(*args)
    if rc != 0:
        msg = lib.vgpt_probe_last_error()
        raise VgptError(f"{name} failed (rc={rc}): {msg.decode() if msg else '?'}")
