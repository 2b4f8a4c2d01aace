"""Latent-space rollout with a persistent paged K/V cache on the GPU.  Host side
and parity definition: videogpt_b200/rollout.py, checked in fp32 against the oracle on CPU by
tests/test_emu_host_numerics.py.  Here, through the C ABI in bf16:
* while the window holds the whole history a round must equal the reference-shaped round
  (LVMPipeline.next_clip_latents on all frames so far) BIT FOR BIT -- cached context rows were
  computed by an earlier, differently sized launch, and every kernel gives a row the same bits
  whatever shares its launch (profiles/r01g_row_partition_invariance.txt);
* with eviction, against the oracle in bf16 on the same device over the full history with the
  history mask (final-latent cosine >= 0.999, BASELINE.json).
Sorts last on purpose: written after this round's GPU minutes were spent."""
import pytest
import torch

from oracle import model_oracle as mo, processor_oracle as po, scheduler_oracle as so
from videogpt_b200 import synth

from helpers import cosine, rel_l2

pytestmark = pytest.mark.gpu
DEV, BF = "cuda", torch.bfloat16


def _pipe(dims):
    from transformers import Phi3Config
    from videogpt_b200 import LVM, LVMPipeline, LVMProcessor
    sd = synth.init_state_dict(dims, seed=0)
    model = LVM(Phi3Config(**dims.phi3_kwargs()), device=DEV)
    model.load_state_dict(sd)
    model.to(BF).eval()
    return LVMPipeline(None, model, LVMProcessor(synth.SingleIdTagTokenizer()), device=torch.device(DEV)), sd


def test_rollout_rounds_equal_reference_rounds_bit_for_bit_without_eviction():
    from videogpt_b200 import LatentRollout
    n0, gen, H, W, steps, rounds = 2, 2, 128, 128, 4, 3
    pipe, _ = _pipe(synth.REDUCED)
    lat = [x.to(DEV, BF) for x in synth.synthetic_latents(n0 + gen * rounds, H, W, seed=5)]
    history = lat[:n0]
    ro = LatentRollout(pipe.model, pipe.processor, gen, max_frame_window=16, num_inference_steps=steps,
                       img_guidance_scale=1.5, prediction_type="x1").start(history)
    outs = [ro.next_clip(initial_noise=lat[n0 + gen * r:n0 + gen * (r + 1)]) for r in range(rounds)]
    torch.cuda.synchronize()
    assert ro.prefilled_frames == n0 + gen * (rounds - 1)
    for r in range(rounds):
        want = pipe.next_clip_latents(history, gen, num_inference_steps=steps, img_guidance_scale=1.5,
                                      prediction_type="x1", initial_noise=lat[n0 + gen * r:n0 + gen * (r + 1)])
        for a, b in zip(outs[r], want):
            assert torch.equal(a, b), f"round {r}: cached-context round differs from the recomputed one"
        history = history + outs[r]


def test_rollout_with_eviction_matches_the_oracle_over_the_full_history():
    from videogpt_b200 import LatentRollout
    from videogpt_b200.rollout import window_start
    n0, gen, window, H, W, steps, rounds = 2, 2, 6, 128, 128, 4, 4
    bl = (H // 16) * (W // 16) + 2
    dims = synth.REDUCED
    pipe, sd = _pipe(dims)
    w = {k: v.to(DEV, BF) for k, v in sd.items()}
    cfg = mo.OracleConfig(hidden_size=dims.hidden_size, intermediate_size=dims.intermediate_size,
                          num_hidden_layers=dims.num_hidden_layers, num_attention_heads=dims.num_attention_heads)
    lat = [x.to(DEV, BF) for x in synth.synthetic_latents(n0 + gen * rounds, H, W, seed=6)]
    history = lat[:n0]
    ro = LatentRollout(pipe.model, pipe.processor, gen, max_frame_window=window, num_inference_steps=steps,
                       img_guidance_scale=1.5, prediction_type="v").start(history)
    ws_of_frame = {}
    for r in range(rounds):
        n_hist = len(history)
        ws = window_start(n_hist, gen, window)
        for f in range(n_hist):
            ws_of_frame.setdefault(f, ws)
        noise = lat[n0 + gen * r:n0 + gen * (r + 1)]
        got = ro.next_clip(initial_noise=noise)
        d = po.frame_block_inputs(n_hist, gen, H, W, True, 1)
        mask = d["attention_mask"].clone()
        frame = torch.arange(mask.shape[-1]) // bl
        for q in range(mask.shape[-1]):
            fq = int(frame[q])
            mask[0, q, frame < (ws if fq >= n_hist else ws_of_frame[fq])] = 0
        mk = dict(input_ids=d["input_ids"].to(DEV), input_img_latents=history, input_image_sizes=d["input_image_sizes"],
                  attention_mask=mask.to(DEV), position_ids=d["position_ids"].to(DEV),
                  denoise_image_sizes=d["denoise_image_sizes"], time_emb_inx=d["time_emb_inx"], img_cfg_scale=1.5,
                  use_img_cfg=True)
        with torch.no_grad():
            want = so.euler_sample([x.clone() for x in noise] * 2,
                                   lambda z, t, **kw: mo.frame_block_forward_with_cfg(w, cfg, z, t, **kw),
                                   mk, num_steps=steps, prediction_type="v")[:gen]
        a, b = torch.cat(got, 0), torch.cat(want, 0)
        assert torch.isfinite(a.float()).all()
        assert cosine(a, b) >= 0.999 and rel_l2(a, b) <= 2e-2, f"round {r}: cos {cosine(a, b)}, rel {rel_l2(a, b)}"
        history = history + got
    assert min(ro._phys) > 0          # pages behind the window went back to the free list
