#!/usr/bin/env python
"""One context prefill + N denoising forwards of a workload, eager (no CUDA graph), bracketed by
cudaProfilerStart/Stop -- the command the ncu passes of profiles/ are taken on:

    ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
        --log-file gpurun_out/launches.csv python tools/profile_step.py --config cfg2
"""
import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import bench  # noqa: E402
from videogpt_b200 import LVMPipeline, LVMProcessor, LVMScheduler, ops, synth  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="cfg2")
    ap.add_argument("--predicts", type=int, default=1)
    ap.add_argument("--no-prefill", action="store_true")
    args = ap.parse_args()
    kind, n_ctx, n_gen, H, W, euler = bench.WORKLOADS[args.config]
    dims = bench._dims(kind)
    dev = torch.device("cuda", 0)
    model = bench.build_model(dims, dev)
    model.use_cuda_graph = False
    pipe = LVMPipeline(None, model, LVMProcessor(synth.SingleIdTagTokenizer()), device=dev)
    lat = [x.to(dev, torch.bfloat16) for x in synth.synthetic_latents(n_ctx + n_gen, H, W)]
    pipe.next_clip_latents(lat[:n_ctx], n_gen, num_inference_steps=2, img_guidance_scale=1.5,
                           prediction_type="x1", initial_noise=lat[n_ctx:])          # warm-up + plan
    e = model.engine()
    e.uniform_t = True           # as the fused sampler loop sets it: one timestep for every latent -> the small MLPs run on one row
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    if not args.no_prefill:
        e.prefill(None)
    for i in range(args.predicts):
        e.t.fill_(0.5)
        e.predict()
        ops.cfg_euler(e.z, e.pred, True, True, 0.5, 0.02, 1.5)
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
    print("profiled: prefill=%s predicts=%d launches/predict=%d" % (not args.no_prefill, args.predicts, e.launches_per_predict + 1))


if __name__ == "__main__":
    main()
