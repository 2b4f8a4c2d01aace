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


def test_launch_breakdown_and_traffic_parsers_read_the_committed_evidence(capsys):
    """tools/launch_breakdown.py on the committed ncu launch list, and bench.py's reader of the committed `--set full`
    summary (the source of roofline.traffic): both run on the GPU box where a parse error would cost a bench line."""
    import launch_breakdown as lb
    lb.main(os.path.join(ROOT, "profiles", "r02h_launches_prefill_plus_1step.csv"))
    out = capsys.readouterr().out
    assert "prefill: 253 launches" in out and "one Euler step: 265 launches" in out
    assert "gemm_bf16_pair_kernel<256, 2>" in out and "attn_pair_tcgen05_kernel<96, 0>" in out
    sys.path.insert(0, ROOT)
    import bench
    traffic, note = bench.gemm_traffic_from_ncu_summary()
    # mean of qkv / o / gate_up / down at M = 2064: between half and twice the algorithmic mean (105 MB)
    assert traffic is not None and 50e6 < traffic < 210e6, (traffic, note)
    assert bench.gemm_traffic_from_ncu_summary(os.path.join(ROOT, "profiles", "does_not_exist.txt"))[0] is None
