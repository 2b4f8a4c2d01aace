"""Parity at BASELINE.json's FULL size and FULL length (configs[1]: Phi-3-mini-class backbone, 4 + 4 frames
256x256, CFG 1.5, 50 Euler steps) against the oracle on the same device, measured by tools/parity_floor.py
(committed run: profiles/r02a_parity_floor.json).

BASELINE states: per-step velocity rel-L2 <= 1e-2 (bf16) against the reference's own PyTorch path, final-latent
cosine >= 0.999.  What is measured here, every run:

  ours  the CUDA path                         A  oracle bf16, GPU eager, default SDPA backend (the gate's reference)
  B     the SAME oracle, bf16, SDPA forced to the MATH backend          F  oracle fp32 (ground truth)

* the bf16-vs-bf16 floor is MEASURED, not inferred: B vs A -- two evaluations of the reference's own bf16 path
  that differ only in the attention backend -- disagree by 2.7e-2 (step 0) ... 8.9e-1 (step 49) in x1 mode and
  4.3e-2 in v mode (the host-CPU evaluation of the same oracle: 2.7e-2 at steps 0 and 1).  The stated 1e-2
  cannot be met at this size by the reference against itself; it is asserted wherever it can hold (reduced
  size, tests/test_model_gpu.py);
* gate 1 (per step): ours-vs-A <= 1.10 x (B-vs-A): the CUDA path is no further from the reference's bf16 path
  than that path is from itself (measured ratio 0.98 ... 1.02 over all 50 steps);
* gate 2 (per step): ours-vs-F <= 1.10 x (A-vs-F): the CUDA path is as close to fp32 as the reference's bf16 path;
* gate 3: final-latent cosine ours-vs-A >= the measured B-vs-A cosine - 1e-4 (x1: 0.99911 vs 0.99912; the stated
  0.999 holds in both modes), and ours-vs-F >= A-vs-F - 1e-4;
* gate 4 (per decoder layer, first forward): hidden-state error of ours vs F <= 1.05 x that of A vs F at EVERY
  layer, generated rows and context rows -- a kernel regression cannot hide inside the end-to-end floor.
Sorts last: the 50-step fp32 oracle takes half a minute per mode."""
import os
import sys

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
STEPS = 50
COS_TOL = 0.999


@pytest.fixture(scope="module")
def case():
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import parity_floor
    from videogpt_b200 import synth
    return parity_floor, parity_floor.Case(synth.FULL_SIZE, 4, 4, 256, 256)


@pytest.mark.timeout(900)
def test_full_size_per_layer_hidden_state_error_is_the_reference_bf16_error(case):
    pf, c = case
    pl = pf.per_layer(c)
    for name in ("gen", "ctx"):
        for i, (a, b) in enumerate(zip(pl[f"{name}_ours_vs_F"], pl[f"{name}_A_vs_F"])):
            assert a <= 1.05 * b + 1e-5, f"{name} rows, layer {i}: ours vs fp32 {a:.3e}, reference bf16 vs fp32 {b:.3e}"


@pytest.mark.timeout(1200)
@pytest.mark.parametrize("pt", ["x1", "v"])
def test_full_size_cfg2_50_steps_velocity_and_final_latents(case, pt):
    pf, c = case
    r = pf.trajectories(c, STEPS, pt, cpu_steps=0)
    for i in range(STEPS):
        e_a, floor_bb = r["vel_ours_vs_A"][i], r["vel_B_vs_A"][i]
        e_f, floor_f = r["vel_ours_vs_F"][i], r["vel_A_vs_F"][i]
        assert e_a <= max(1e-2, 1.10 * floor_bb), f"step {i}: ours vs bf16 oracle {e_a:.3e}, bf16-vs-bf16 floor {floor_bb:.3e}"
        assert e_f <= 1.10 * floor_f + 1e-3, f"step {i}: ours vs fp32 {e_f:.3e}, reference bf16 vs fp32 {floor_f:.3e}"
    fc = r["final_cos"]
    assert fc["ours_vs_A"] >= min(COS_TOL, fc["B_vs_A"]) - 1e-4, fc
    assert fc["ours_vs_A"] >= COS_TOL, fc          # holds at this size in both modes (x1: 0.99911)
    assert fc["ours_vs_F"] >= fc["A_vs_F"] - 1e-4, fc
