"""Parity at BASELINE.json's FULL size (configs[1]: Phi-3-mini-class backbone, 4 + 4 frames 256x256,
CFG) against the oracle on the same device, through the same measurement as tools/parity_report.py
(profiles/r01a_parity.json): per-step velocity of ours vs the bf16 oracle, ours vs the fp32 oracle,
and the bf16 oracle's own distance to fp32 (the noise floor of the reference's path).

At this size two independent bf16 evaluations cannot agree to 1e-2 -- the reference against itself
included (floor 2.7e-2 in x1 mode at step 0, 4.3e-2 in v mode) -- so the gate is
BASELINE's 1e-2 wherever the floor allows it and 1.3 x the floor elsewhere, plus: ours is as close to
fp32 as the reference's own bf16 path is (1.15 x floor), and the final-latent cosine holds.
Sorts last: a full-size run takes a minute or two."""
import os
import sys

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
STEPS = 4


@pytest.mark.parametrize("pt", ["x1", "v"])
def test_full_size_cfg2_velocity_and_final_latents_match_the_oracle(pt):
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import parity_report
    from videogpt_b200 import synth
    r = parity_report.run(synth.FULL_SIZE, 4, 4, 256, 256, STEPS, pt)
    for i in range(STEPS):
        floor = r["velocity_rel_l2_oracle_bf16_vs_fp32"][i]
        err = r["velocity_rel_l2_ours_vs_oracle_bf16"][i]
        err32 = r["velocity_rel_l2_ours_vs_oracle_fp32"][i]
        assert err <= max(1e-2, 1.3 * floor), f"step {i}: velocity rel-L2 vs bf16 oracle {err:.3e} (floor {floor:.3e})"
        assert err32 <= 1.15 * floor + 1e-3, f"step {i}: vs fp32 oracle {err32:.3e} (reference bf16 {floor:.3e})"
    cos, cos_floor = r["final_cosine_ours_vs_oracle_bf16"], r["final_cosine_oracle_bf16_vs_fp32"]
    assert cos >= min(0.999, cos_floor - 2e-4), (cos, cos_floor)
    assert r["final_cosine_ours_vs_oracle_fp32"] >= cos_floor - 2e-4
