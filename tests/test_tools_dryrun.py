"""Dry runs of the GPU tools' host logic on the CPU emulation (a tool that crashes on the GPU box costs
GPU minutes): tools/parity_floor.py with the reduced backbone, all evaluations in fp32 on the CPU."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))


def test_parity_floor_tool_dry_run(emu, monkeypatch):
    import parity_floor as pf
    from videogpt_b200 import synth
    monkeypatch.setattr(pf, "DEV", "cpu")
    monkeypatch.setattr(pf, "BF", torch.float32)
    case = pf.Case(synth.REDUCED, 2, 2, 64, 64)
    pl = pf.per_layer(case)
    nl = synth.REDUCED.num_hidden_layers
    assert len(pl["gen_ours_vs_F"]) == nl and len(pl["ctx_ours_vs_F"]) == nl - 1
    # fp32 emulation of the product path vs the fp32 oracle: rounding only
    assert max(pl["gen_ours_vs_F"]) < 1e-4 and max(pl["ctx_ours_vs_F"]) < 1e-4
    r = pf.trajectories(case, 3, "x1", cpu_steps=2)
    assert len(r["vel_ours_vs_A"]) == 3 and len(r["vel_C_vs_A"]) == 2
    assert max(r["vel_ours_vs_A"]) < 1e-3 and r["final_cos"]["ours_vs_F"] > 0.99999
    r = pf.trajectories(case, 2, "v", cpu_steps=0)
    assert max(r["vel_ours_vs_F"]) < 1e-3
