"""CPU-side checks: the C-ABI library loads and exports every symbol include/vgpt_b200.h declares
(no compute calls without a GPU), the host-side plan builder, and the no-fallback contract."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

from oracle import processor_oracle as po
from videogpt_b200 import _lib, engine as eng, ops, synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    text = open(os.path.join(ROOT, "include", "vgpt_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    out = {}
    for m in re.finditer(r"(?:int|size_t|const char\*)\s+(vgpt_\w+)\s*\(([^;]*?)\)\s*;", text, flags=re.S):
        args = m.group(2).strip()
        out[m.group(1)] = 0 if args in ("", "void") else len(args.split(","))
    return out


def test_library_exports_every_declared_symbol():
    decl = _declared()
    assert len(decl) >= 16
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name, nargs in decl.items():
        assert hasattr(lib, name), f"{name} declared in include/vgpt_b200.h but not exported"
        assert name in _lib.SIGNATURES, f"{name} has no ctypes binding"
        assert len(_lib.SIGNATURES[name]) == nargs, f"{name}: binding has {len(_lib.SIGNATURES[name])} args, header {nargs}"
    assert set(_lib.SIGNATURES) == set(decl)
    assert _lib.load().vgpt_abi_version() == 1


def test_argument_errors_surface_as_exceptions_without_a_gpu():
    with pytest.raises(_lib.VgptError, match="null pointer"):
        _lib.call("vgpt_rmsnorm", None, None, None, 1, 8, 1e-5, None)
    with pytest.raises(_lib.VgptError, match="multiple of 64"):
        buf = ctypes.create_string_buffer(64)
        p = ctypes.cast(buf, ctypes.c_void_p)
        _lib.call("vgpt_gemm_bf16", p, p, p, None, 8, 64, 100, 104, 64, 0, 0, 0, None)


def test_no_cpu_fallback():
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ops.gemm(torch.zeros(8, 64, dtype=torch.bfloat16), torch.zeros(64, 64, dtype=torch.bfloat16))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ops.rmsnorm(torch.zeros(2, 8, dtype=torch.bfloat16), torch.ones(8, dtype=torch.bfloat16), 1e-5)
    if not torch.cuda.is_available():
        from transformers import Phi3Config
        from videogpt_b200 import LVM, LVMScheduler
        m = LVM(Phi3Config(**synth.REDUCED.phi3_kwargs()), device="cpu", materialize_pos_embed=False)
        with pytest.raises(RuntimeError, match="no CPU fallback"):
            m.engine()
        with pytest.raises(RuntimeError, match="no CPU fallback"):
            LVMScheduler(2)(torch.zeros(2, 4, 8, 8), lambda *a, **k: None, {"use_img_cfg": False})


def test_missing_library_is_loud(tmp_path, monkeypatch):
    monkeypatch.setattr(_lib, "_lib", None)
    with pytest.raises(ImportError, match="no CPU fallback"):
        _lib.load(str(tmp_path / "libvgpt_b200.so"))
    monkeypatch.setattr(_lib, "_lib", None)


CASES = [(4, 4, 256, 256, 1), (3, 4, 176, 320, 8), (5, 2, 64, 96, 8), (1, 1, 64, 64, 1), (8, 4, 176, 320, 4)]


@pytest.mark.parametrize("case", CASES)
def test_plan_codes_reproduce_the_reference_mask(case):
    n_ctx, n_gen, H, W, sp = case
    d = po.frame_block_inputs(n_ctx, n_gen, H, W, True, sp)
    specs, n_lat, n_c = eng.frame_block_specs(d["input_ids"], d["position_ids"], d["input_image_sizes"],
                                              d["denoise_image_sizes"], d["time_emb_inx"])
    L = d["input_ids"].shape[1]
    for b, s in enumerate(specs):
        pad = L - (s.n_prefix + s.n_active)
        assert np.array_equal(eng.codes_dense_mask(s.codes, pad), d["attention_mask"][b].numpy())
        assert np.array_equal(s.positions, d["position_ids"][b, pad:].numpy())
    assert n_lat == 2 * n_gen and n_c == n_ctx
    assert specs[1].n_prefix == 0 and specs[0].n_prefix == n_ctx * (H * W // 256 + 2)


def test_plan_arrays():
    n_ctx, n_gen, H, W = 4, 4, 256, 256
    d = po.frame_block_inputs(n_ctx, n_gen, H, W, True, 1)
    specs, n_lat, n_c = eng.frame_block_specs(d["input_ids"], d["position_ids"], d["input_image_sizes"],
                                              d["denoise_image_sizes"], d["time_emb_inx"])
    plan = eng.build_plan(specs, n_lat, n_c, H // 8, W // 8, "cpu")
    assert plan.total_pages == 17 + 9 and plan.page_table.shape == (2, 17)
    assert plan.prefix.rows == 1032 and plan.step.rows == 2064
    assert plan.step.seqs.tolist() == [[0, 1032, 2064, 0], [1032, 1032, 1032, 0]]
    assert plan.prefix.seqs.tolist() == [[0, 1032, 1032, 0], [1032, 0, 0, 0]]
    # every (sequence, logical position) owns a distinct KV slot
    slots = torch.cat([plan.prefix.row_slot, plan.step.row_slot]).tolist()
    assert len(set(slots)) == len(slots) == 3096 and max(slots) < plan.total_pages * 128
    # RoPE positions: cond generated rows continue after the context, uncond restart at 0 (quirk q9)
    assert plan.step.row_pos[:3].tolist() == [1032, 1033, 1034] and plan.step.row_pos[1032:1035].tolist() == [0, 1, 2]
    # rows of the image tokens of each latent (cond 0..3, uncond 4..7)
    assert plan.lat_row0.tolist() == [2 + 258 * j for j in range(4)] + [1032 + 2 + 258 * j for j in range(4)]
    # assembly kinds of a generated block: tag, time slot, 256 noisy patches
    assert plan.step.kind[:4].tolist() == [ops.ROW_TOKEN, ops.ROW_TIME, ops.ROW_NOISY_PATCH, ops.ROW_NOISY_PATCH]
    assert plan.step.arg_a[:3].tolist() == [32003, 0, 0] and plan.step.arg_b[2:5].tolist() == [0, 1, 2]
    assert plan.prefix.kind[:2].tolist() == [ops.ROW_TOKEN, ops.ROW_CONTEXT_PATCH] and int(plan.prefix.kind[257]) == ops.ROW_TOKEN
    # per-tile (min, max) of the key codes
    mm = plan.k_tile_minmax[0]
    codes = specs[0].codes
    for t in range(33):
        seg = codes[t * 64:(t + 1) * 64]
        assert mm[t].tolist() == [int(seg.min()), int(seg.max())]


def test_prefix_caching_is_refused_when_unsound():
    s = eng.SequenceSpec(n_prefix=2, n_active=2, positions=np.arange(4, dtype=np.int32),
                         codes=np.array([0, 5, 5, 6], np.int32), kinds=np.zeros(4, np.int32),
                         arg_a=np.zeros(4, np.int32), arg_b=np.zeros(4, np.int32))
    with pytest.raises(ValueError, match="prefix"):
        eng.build_plan([s], 0, 0, 8, 8, "cpu")


def test_frame_block_specs_reject_malformed_index_dicts():
    d = po.frame_block_inputs(2, 2, 64, 64, True, 1)
    bad = {k: list(v) for k, v in d["time_emb_inx"].items()}
    bad[0] = [x + 1 for x in bad[0]]
    with pytest.raises(ValueError):
        eng.frame_block_specs(d["input_ids"], d["position_ids"], d["input_image_sizes"], d["denoise_image_sizes"], bad)
    with pytest.raises(ValueError):
        eng.frame_block_specs(d["input_ids"], d["position_ids"], d["input_image_sizes"], {0: [], 1: []}, d["time_emb_inx"])


def test_scheduler_sigma_and_scalars():
    from videogpt_b200 import LVMScheduler
    from oracle import scheduler_oracle as so
    for steps, shift in ((4, 1), (50, 1), (50, 3.0)):
        s = LVMScheduler(steps, shift)
        assert torch.equal(s.sigma, so.sigma_grid(steps, shift))
        oms, ds = s._scalars(steps - 1)
        assert oms == float(1.0 - s.sigma[steps - 1]) and ds == float(s.sigma[steps] - s.sigma[steps - 1])


@pytest.mark.parametrize("case", [(2, 64, 64, 1), (4, 256, 256, 1), (3, 176, 320, 8), (1, 64, 96, 4)])
def test_single_frame_codes_reproduce_the_reference_mask(case):
    n_ctx, H, W, sp = case
    d = po.single_frame_inputs(n_ctx, H, W, True, sp)
    n_tok = H * W // 256
    specs, n_lat, n_c = eng.single_frame_specs(d["input_ids"], d["position_ids"], d["input_image_sizes"], n_tok)
    L = d["position_ids"].shape[1]
    assert n_lat == 2 and n_c == n_ctx
    for b, s in enumerate(specs):
        pad = L - (s.n_prefix + s.n_active)
        assert np.array_equal(eng.codes_dense_mask(s.codes, pad), d["attention_mask"][b].numpy().astype(bool))
        assert np.array_equal(s.positions, d["position_ids"][b, pad:].numpy())
    plan = eng.build_plan(specs, n_lat, n_c, H // 8, W // 8, "cpu")
    assert plan.step.rows == 2 * (n_tok + 1) and plan.prefix.rows == specs[0].n_prefix + 1


@pytest.mark.parametrize("case", [(3, 2, 64, 64, 1), (2, 3, 64, 96, 8)])
def test_codes_from_mask_recovers_block_causal_masks(case):
    from videogpt_b200.transform import codes_from_mask
    n_ctx, n_gen, H, W, sp = case
    for d in (po.frame_block_inputs(n_ctx, n_gen, H, W, True, sp), po.single_frame_inputs(n_ctx, H, W, True, sp)):
        for b in range(d["attention_mask"].shape[0]):
            m = d["attention_mask"][b].bool()
            qc, kc = codes_from_mask(m)
            assert torch.equal(qc[:, None] >= kc[None, :], m)
    with pytest.raises(ValueError):
        codes_from_mask(torch.eye(5, dtype=torch.bool).flip(0))


def test_bench_reference_arm_prints_the_contract_line_on_cpu():
    """`bench.py --impl reference` (the reference's path on the host cores: oracle port) needs no GPU;
    its one JSON line must carry the keys the driver's ratio is computed from."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--config", "cfg1",
                        "--steps", "1", "--warmup", "0"], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["metric"] == "next_clip_tokens_per_s" and line["unit"] == "tokens/s"
    assert line["value"] > 0 and line["higher_is_better"] is True and line["config"]["workload"] == "cfg1"
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1 and line["cpu_baseline"]["sample"]
    assert line["e2e"] == {"value": line["value"], "unit": "tokens/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_scheduler_sigma_grid_matches_the_reference():
    import torch
    from videogpt_b200 import LVMScheduler
    assert torch.allclose(LVMScheduler(num_steps=2).sigma, torch.tensor([0.0, 0.5, 1.0]))

