#!/usr/bin/env python
"""Per-round time of an autoregressive latent rollout (CUDA events around every round), persistent paged K/V cache
(`rollout.LatentRollout`: every context frame prefilled once) against the reference flow (every round recomputes its
whole window).  bench.py --rollout R times R rounds as one step, restarts included; this shows the steady state:

    python tools/rollout_bench.py [rounds=8] [config=cfg3]
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import bench  # noqa: E402
from videogpt_b200 import LVMPipeline, LVMProcessor, synth  # noqa: E402
from videogpt_b200.rollout import LatentRollout  # noqa: E402


def main():
    rounds = int(sys.argv[1]) if len(sys.argv) > 1 else 8
    config = sys.argv[2] if len(sys.argv) > 2 else "cfg3"
    kind, n_ctx, n_gen, H, W, euler = bench.WORKLOADS[config]
    dev = torch.device("cuda", 0)
    model = bench.build_model(bench._dims(kind), dev)
    pipe = LVMPipeline(None, model, LVMProcessor(synth.SingleIdTagTokenizer()), device=dev)
    lat = [x.to(dev, torch.bfloat16) for x in synth.synthetic_latents(n_ctx + n_gen, H, W, seed=42)]
    ctx = lat[:n_ctx]
    window = n_ctx + n_gen

    def timed(fn):
        s, t = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); out = fn(); t.record(); torch.cuda.synchronize()
        return s.elapsed_time(t), out

    for label in ("warm-up", "timed"):
        ro = LatentRollout(model, pipe.processor, n_gen, window, euler, bench.GUIDANCE, True, 1.0, "x1", 0.0,
                           rounds_hint=rounds).start([x.clone() for x in ctx])
        pers = [timed(lambda: ro.next_clip(seed=7))[0] for _ in range(rounds)]
        frames, reco = [x.clone() for x in ctx], []
        for _ in range(rounds):
            if len(frames) + n_gen > window:
                frames = frames[n_gen + len(frames) - window:]
            ms, out = timed(lambda: pipe.next_clip_latents(frames, n_gen, num_inference_steps=euler,
                                                           img_guidance_scale=bench.GUIDANCE, prediction_type="x1", seed=7))
            reco.append(ms)
            frames = frames + out
        if label == "timed":
            fmt = lambda xs: " ".join(f"{x:7.1f}" for x in xs)
            print(f"{config}: {n_ctx} context + {n_gen} generated frames {H}x{W}, {euler} Euler steps, window {window} frames; ms per round")
            print(f"persistent K/V cache : {fmt(pers)}   (steady state, rounds 3+: {sum(pers[2:]) / max(len(pers) - 2, 1):7.1f})")
            print(f"recompute (reference): {fmt(reco)}   (steady state, rounds 3+: {sum(reco[2:]) / max(len(reco) - 2, 1):7.1f})")


if __name__ == "__main__":
    main()
