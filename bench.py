#!/usr/bin/env python
"""Benchmark of the next-clip denoising hot path (BASELINE.json: next-clip latency / tokens/s).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config cfg2|cfg3|cfg4|cfg5|cfg1]

One "step" = one next-clip prediction of the named workload: context prefill + all Euler
steps with CFG (50 at full size), synthetic latents, random-init weights.  Prints ONE JSON line
(rank 0).  See DESIGN.md "Measurement" for how every field is obtained.

* ``value``   : tokens/s = (1+cfg) * T_gen * euler_steps * clips / time, latents resident in HBM,
                timed with CUDA events, max over ranks.
* ``e2e``     : same metric through ``LVMPipeline.next_clip_latents`` with HOST latents in pinned
                memory and the generated latents read back to the host, inside the timed region.
* ``roofline``: the tcgen05 GEMM (dominant kernel) timed live with CUDA events on the launch
                stream over all layers' weights (7.2 GB > L2), algorithmic FLOPs / duration.
* ``gpu_eager_baseline`` (N = 1, full size): the same oracle restatement run in bf16 eager on the
                same GPU and weights, one Euler step -- the reference's path as written on this
                hardware (SURVEY.md 8(d) "GPU reference baseline").  A measured baseline only:
                nothing in ``videogpt_b200`` imports ``oracle``.
* ``cpu_baseline`` / ``--impl reference``: the oracle restatement of the reference's own
                PyTorch path (no cache, padded unconditional row, dense mask) on the host cores:
                one COMPLETE Euler step (all layers, both CFG rows) timed, x the number of steps,
                labelled as such.
N > 1 (torchrun): --parallelism dp = independent videos data-parallel across ranks (weak scaling,
no collective; the default); cfg = CFG branches on rank pairs; sp = sequence parallel, groups of
--sp ranks share one video (rows sharded, K/V stored into the peers over NVLink: strong scaling
of one video inside a group, weak across groups).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (dims, n_ctx, n_gen, height, width, euler_steps)
    "cfg2": ("full", 4, 4, 256, 256, 50),   # BASELINE.json configs[1]: the config the metric is quoted on
    "cfg3": ("full", 32, 4, 256, 256, 50),  # configs[2]: 8 context clips, KV cache (sequence parallel at N>1)
    "cfg5": ("full", 4, 4, 512, 512, 50),   # configs[4]
    "cfg1": ("reduced", 4, 4, 256, 256, 4),  # configs[0] (the reference's CPU-runnable case)
    "cfg4": ("full", 4, 4, 256, 256, 50),   # configs[3]: 32 videos x CFG branches, data-parallel over the ranks
}
BATCH_VIDEOS = {"cfg4": 32}                 # total videos of the job (split over the ranks: strong scaling)
GUIDANCE = 1.5


NCU_GEMM_SUMMARY = os.path.join("profiles", "r02h_gemm_pair_ncu.txt")     # tools/ncu_summary.py of the --set full capture


def gemm_traffic_from_ncu_summary(path=None):
    """(bytes, note): dram__bytes_read.sum + dram__bytes_write.sum per launch, averaged over the GEMM launches of the
    committed ncu summary (cfg2 step: qkv, o, gate_up, down at M = 2064; algorithmic: 107 / 57 / 147 / 110 MB);
    (None, why) when the file is missing or unreadable."""
    path = path or os.path.join(ROOT, NCU_GEMM_SUMMARY)
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    try:
        per_kernel, cur = [], None
        for line in open(path):
            if line.startswith("kernel:"):
                cur = 0.0 if "gemm_bf16" in line else None
                if cur is not None:
                    per_kernel.append(0.0)
            elif cur is not None and ("dram__bytes_read.sum" in line or "dram__bytes_write.sum" in line):
                _, val, unit = line.split()[:3]
                per_kernel[-1] += float(val.replace(",", "")) * scale[unit]
        if not per_kernel:
            return None, f"no GEMM launch in {NCU_GEMM_SUMMARY}"
        return sum(per_kernel) / len(per_kernel), (f"dram__bytes_read + dram__bytes_write per launch, mean of the {len(per_kernel)} "
                                                   f"GEMM launches of a layer in {NCU_GEMM_SUMMARY} (ncu --set full, M = 2064)")
    except Exception as exc:          # the number is evidence, not a dependency of the run
        return None, f"{NCU_GEMM_SUMMARY}: {type(exc).__name__}"


def _dims(kind):
    from videogpt_b200 import synth
    return synth.FULL_SIZE if kind == "full" else synth.REDUCED


def algorithmic_flops(dims, n_ctx, n_gen, block, euler_steps):
    """per clip, context K/V cached, uncond row unpadded."""
    h, i, L = dims.hidden_size, dims.intermediate_size, dims.num_hidden_layers
    per_tok = 2 * (4 * h * h + 3 * h * i) * L
    t_ctx, t_gen = n_ctx * block, n_gen * block
    step = 2 * t_gen * per_tok + L * 4 * h * (t_gen * (t_ctx + t_gen) + t_gen * t_gen)
    prefill = t_ctx * per_tok + L * 4 * h * block * block * sum(range(1, n_ctx + 1))
    return step, prefill, step * euler_steps + prefill


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc is not None:
            self.proc.terminate()
        sm = sorted(int(r[0]) for r in self.rows if r and r[0].isdigit())
        mx = [int(r[1]) for r in self.rows if len(r) > 1 and r[1].isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for k, n in enumerate(names) if any(len(r) > 2 + k and r[2 + k].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


def build_model(dims, device):
    """Random-init full model directly on the GPU (no checkpoint: no network), final layer re-drawn
    N(0, 0.02) because the reference zero-initialises it (LVM/model.py:241-244)."""
    from transformers import Phi3Config
    from videogpt_b200 import LVM
    torch.manual_seed(0)
    model = LVM(Phi3Config(**dims.phi3_kwargs()), device=device, materialize_pos_embed=False)
    with torch.no_grad():
        for lin in (model.final_layer.linear, getattr(model.final_layer.adaLN_modulation, "1")):
            lin.weight.normal_(std=0.02)
        for n, p in model.named_parameters():
            if n.endswith("layernorm.weight") or n == "llm.norm.weight":
                p.uniform_(0.9, 1.1)
    model.to(torch.bfloat16).eval()
    return model


def gemm_roofline(model, rows, reps=3):
    """Time every projection GEMM of one Euler step (4 per layer x all layers, real weights) with
    CUDA events on the launch stream; returns (avg seconds per launch, flops per launch, launches)."""
    from videogpt_b200 import ops
    e = model.engine()
    h, i = e.hs, e.inter
    x = torch.randn(rows, h, device=e.device).to(torch.bfloat16)
    xi = torch.randn(rows, i, device=e.device).to(torch.bfloat16)
    qkv = torch.empty(rows, 3 * h, device=e.device, dtype=torch.bfloat16)
    hid = torch.zeros(rows, h, device=e.device, dtype=torch.bfloat16)
    mh = torch.empty(rows, i, device=e.device, dtype=torch.bfloat16)

    def one_pass():
        for lw in e.w.layers:
            ops.gemm(x, lw["qkv"], out=qkv)
            ops.gemm(x, lw["o"], out=hid, residual=hid, epilogue=ops.EPI_RESIDUAL)
            ops.gemm(x, lw["gate_up"], out=mh, epilogue=ops.EPI_SWIGLU)
            ops.gemm(xi, lw["down"], out=hid, residual=hid, epilogue=ops.EPI_RESIDUAL)

    one_pass()
    torch.cuda.synchronize()
    s, t = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(reps):
        one_pass()
    t.record()
    torch.cuda.synchronize()
    launches = reps * 4 * len(e.w.layers)
    flops_per_layer = 2 * rows * (4 * h * h + 3 * h * i)
    return s.elapsed_time(t) * 1e-3 / launches, flops_per_layer / 4.0, launches


_CPU_SAMPLE_CACHE = {}


def cpu_reference_sample(dims, n_ctx, n_gen, height, width, euler_steps, layers_distinct, threads, layers_timed=None):
    """The reference's own PyTorch path (oracle restatement: no cache, padded uncond row, dense mask) on the host
    cores: ONE COMPLETE Euler step -- every decoder layer, both CFG rows, the full padded length -- timed, then
    multiplied by euler_steps ("time >= 1 complete Euler step ... report x50 extrapolation").  Only
    `layers_distinct` layers' worth of random weights are generated (initialising 3.6 B parameters would cost more than
    the step); the layers of the model cycle through them, which changes neither the arithmetic nor -- at 226 MB per
    layer -- what the caches see.  `layers_timed` < all layers (a box with few cores and a caller that wants many
    samples: `cpu_sample_layers`) times a step of that many layers and scales it to all of them, and says so.
    Returns (tokens/s, description, seconds of the timed step)."""
    from oracle import model_oracle as mo, processor_oracle as po
    from videogpt_b200 import synth
    torch.set_num_threads(threads)
    L = dims.num_hidden_layers
    layers_distinct = min(layers_distinct, L)
    layers_timed = L if layers_timed is None else max(1, min(L, layers_timed))
    dtype = torch.bfloat16 if dims.hidden_size >= 1024 else torch.float32
    key = (dims.hidden_size, dims.intermediate_size, L, layers_distinct, n_ctx, n_gen, height, width)
    if key not in _CPU_SAMPLE_CACHE:
        _CPU_SAMPLE_CACHE.clear()
        small = synth.BackboneDims(hidden_size=dims.hidden_size, intermediate_size=dims.intermediate_size,
                                   num_hidden_layers=layers_distinct, num_attention_heads=dims.num_attention_heads,
                                   vocab_size=dims.vocab_size)
        sd = synth.init_state_dict(small, seed=0, dtype=dtype, with_pos_embed=False)
        sd["pos_embed"] = torch.zeros(1, dims.pos_embed_max_size ** 2, dims.hidden_size, dtype=dtype)
        for n in range(layers_distinct, L):          # layer n shares the tensors of layer n % layers_distinct
            src = f"llm.layers.{n % layers_distinct}."
            for k in [k for k in sd if k.startswith(src)]:
                sd[f"llm.layers.{n}." + k[len(src):]] = sd[k]
        d = po.frame_block_inputs(n_ctx, n_gen, height, width, True, 1)
        lat = [x.to(dtype) for x in synth.synthetic_latents(n_ctx + n_gen, height, width, seed=42)]
        _CPU_SAMPLE_CACHE[key] = (sd, d, lat)
    sd, d, lat = _CPU_SAMPLE_CACHE[key]
    cfg = mo.OracleConfig(hidden_size=dims.hidden_size, intermediate_size=dims.intermediate_size,
                          num_hidden_layers=layers_timed, num_attention_heads=dims.num_attention_heads)
    args = (d["input_ids"], lat[:n_ctx], d["input_image_sizes"], d["attention_mask"], d["position_ids"],
            d["denoise_image_sizes"], d["time_emb_inx"])
    z = lat[n_ctx:] * 2
    t = torch.full((len(z),), 0.5)
    with torch.no_grad():
        t0 = time.perf_counter()
        mo.frame_block_forward(sd, cfg, z, t, *args)
        dt = time.perf_counter() - t0
    per_step = dt * L / layers_timed
    block = height * width // 256 + 2
    tokens = 2 * n_gen * block * euler_steps
    what = (f"1 complete Euler step (all {L} layers" if layers_timed == L else
            f"1 Euler step x {layers_timed} of {L} layers (scaled to {L}")
    desc = (f"oracle port of the reference path ({str(dtype).split('.')[-1]}): {what}, both CFG rows, "
            f"L={d['input_ids'].shape[1]}; weights of {layers_distinct} distinct layers cycled) timed ({dt:.2f} s) "
            f"x {euler_steps} steps")
    return tokens / (per_step * euler_steps), desc, dt


def cpu_sample_layers(dims, n_ctx, n_gen, height, width, euler_steps, layers_distinct, threads, samples, budget_s):
    """How many layers a timed CPU step can have so that `samples` of them fit in `budget_s` seconds on THIS box (all of
    them on the GPU boxes seen so far: 16-24 cores, ~0.16 s per layer at cfg2); calibrated with one short step."""
    L = dims.num_hidden_layers
    probe = min(layers_distinct, L)
    _, _, dt = cpu_reference_sample(dims, n_ctx, n_gen, height, width, euler_steps, layers_distinct, threads, probe)
    per_layer = dt / probe
    return max(probe, min(L, int(budget_s / max(samples, 1) / max(per_layer, 1e-9))))


def gpu_eager_oracle(model, dims, n_ctx, n_gen, height, width, dev, reps=3):
    """The reference's own PyTorch path (oracle restatement: every context row recomputed, padded
    unconditional row, dense additive mask, cuBLAS + SDPA, bf16 eager) on the SAME GPU with the same
    weights: seconds per Euler step.  SURVEY.md 8(d) names this, not the CPU number, as the kernel
    bar.  A baseline that is measured, never the product path."""
    from oracle import model_oracle as mo, processor_oracle as po
    from videogpt_b200 import synth
    bf = torch.bfloat16
    w = dict(model.state_dict())
    if w.get("pos_embed") is None:
        w["pos_embed"] = torch.zeros(1, dims.pos_embed_max_size ** 2, dims.hidden_size, device=dev, dtype=bf)
    cfg = mo.OracleConfig(hidden_size=dims.hidden_size, intermediate_size=dims.intermediate_size,
                          num_hidden_layers=dims.num_hidden_layers, num_attention_heads=dims.num_attention_heads)
    d = po.frame_block_inputs(n_ctx, n_gen, height, width, True, 1)
    lat = [x.to(dev, bf) for x in synth.synthetic_latents(n_ctx + n_gen, height, width, seed=42)]
    args = (d["input_ids"].to(dev), lat[:n_ctx], d["input_image_sizes"], d["attention_mask"].to(dev),
            d["position_ids"].to(dev), d["denoise_image_sizes"], d["time_emb_inx"])
    z = lat[n_ctx:] * 2
    t = torch.full((len(z),), 0.5, device=dev)
    with torch.no_grad():
        mo.frame_block_forward(w, cfg, z, t, *args)
        torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(reps):
            mo.frame_block_forward(w, cfg, z, t, *args)
        e.record()
        torch.cuda.synchronize()
    return s.elapsed_time(e) * 1e-3 / reps


def run_reference(args, rank, world):
    if rank != 0:
        return
    kind, n_ctx, n_gen, H, W, euler = WORKLOADS[args.config]
    dims = _dims(kind)
    threads = os.cpu_count() or 1
    layers_distinct = 4 if kind == "full" else dims.num_hidden_layers
    # each step = one complete Euler step of the reference path -- unless warmup + steps of them would not fit in a few
    # minutes on this box's cores; then fewer layers per step, scaled and labelled (the contract: a bounded sample)
    layers_timed = cpu_sample_layers(dims, n_ctx, n_gen, H, W, euler, layers_distinct, threads, args.warmup + args.steps, 150.0)
    vals, times = [], []
    for i in range(args.warmup + args.steps):
        v, desc, dt = cpu_reference_sample(dims, n_ctx, n_gen, H, W, euler, layers_distinct, threads, layers_timed)
        if i >= args.warmup:
            vals.append(v); times.append(dt)
    value = sum(vals) / len(vals)
    block = H * W // 256 + 2
    line = {"impl": "reference", "metric": "next_clip_tokens_per_s", "value": value, "unit": "tokens/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * (2 * n_gen * block * euler) / value, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16" if kind == "full" else "f32", "data": "synthetic",
            "config": {"workload": args.config, "context_frames": n_ctx, "generated_frames": n_gen,
                       "height": H, "width": W, "euler_steps": euler, "cfg": True},
            "cpu_baseline": {"value": value, "unit": "tokens/s", "cores": threads, "kind": "port", "sample": desc},
            "e2e": {"value": value, "unit": "tokens/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def run_ours(args, rank, world, local_rank):
    import torch.distributed as dist
    from videogpt_b200 import LVMPipeline, LVMProcessor, LVMScheduler, synth
    from videogpt_b200.synth import SingleIdTagTokenizer as FakeTokenizer
    kind, n_ctx, n_gen, H, W, euler = WORKLOADS[args.config]
    dims = _dims(kind)
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    sp_size = 1
    cfg_split = args.parallelism == "cfg" and world > 1
    if args.parallelism == "sp" and world > 1:
        from videogpt_b200 import parallel_states
        sp_size = args.sp or world
        parallel_states.initialize_sequence_parallel_state(sp_size)   # the reference's SP switch
    elif cfg_split:
        from videogpt_b200 import parallel_states
        sp_size = 2                                                   # pairs of ranks, one CFG branch each
        parallel_states.initialize_cfg_branch_parallel_state()
    model = build_model(dims, dev)
    pipe = LVMPipeline(None, model, LVMProcessor(FakeTokenizer()), device=dev)
    block = H * W // 256 + 2
    lat = synth.synthetic_latents(n_ctx + n_gen, H, W, seed=42 + rank // sp_size)
    ctx_host = [x.to(torch.bfloat16).pin_memory() for x in lat[:n_ctx]]
    noise_host = [x.to(torch.bfloat16).pin_memory() for x in lat[n_ctx:]]
    ctx_dev = [x.to(dev) for x in ctx_host]
    noise_dev = [x.to(dev) for x in noise_host]
    kw = dict(num_inference_steps=euler, img_guidance_scale=GUIDANCE, prediction_type="x1")

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # configs[3]: this rank's share of the job's videos, `--batch` videos per pass through the engine
    # (all rows of those videos and both CFG branches in one [M, hidden] matrix)
    vids_rank = 0
    if args.config in BATCH_VIDEOS:
        total = args.videos or BATCH_VIDEOS[args.config]
        if args.parallelism != "dp" or total % world or (total // world) % args.batch:
            raise SystemExit(f"{args.config}: {total} videos need --parallelism dp, a world size dividing {total} "
                             f"and --batch dividing {total}//world")
        vids_rank = total // world
        vlat = [synth.synthetic_latents(n_ctx + n_gen, H, W, seed=1000 + rank * vids_rank + v) for v in range(vids_rank)]
        vctx_host = [[x.to(torch.bfloat16).pin_memory() for x in l[:n_ctx]] for l in vlat]
        vnoise_host = [[x.to(torch.bfloat16).pin_memory() for x in l[n_ctx:]] for l in vlat]
        vctx_dev = [[x.to(dev) for x in c] for c in vctx_host]
        vnoise_dev = [[x.to(dev) for x in c] for c in vnoise_host]

    def batch_pass(ctxs, noises):
        out = []
        for i in range(0, vids_rank, args.batch):
            out += pipe.next_clip_latents_batch(ctxs[i:i + args.batch], n_gen, initial_noise=noises[i:i + args.batch], **kw)
        return out

    rounds = max(args.rollout, 1)
    roll_kw = dict(num_inference_steps=euler, img_guidance_scale=GUIDANCE, prediction_type="x1", seed=7,
                   max_frame_window=n_ctx + n_gen, persistent_cache=not args.recompute)

    def clip_device():
        if args.rollout:     # `rounds` clips autoregressively, latents and K/V carried forward
            return pipe.rollout_latents([x.clone() for x in ctx_dev], [n_gen] * rounds, **roll_kw)
        if vids_rank:
            return batch_pass([[x.clone() for x in c] for c in vctx_dev], vnoise_dev)
        # a fresh context tensor list every clip => the engine re-runs the prefill (as a new clip would)
        return pipe.next_clip_latents([x.clone() for x in ctx_dev], n_gen, initial_noise=noise_dev, **kw)

    def clip_host():
        if args.rollout:
            return [x.to("cpu") for x in pipe.rollout_latents(ctx_host, [n_gen] * rounds, **roll_kw)]
        if vids_rank:
            return [[x.to("cpu") for x in v] for v in batch_pass(vctx_host, vnoise_host)]
        out = pipe.next_clip_latents(ctx_host, n_gen, initial_noise=noise_host, **kw)
        return [x.to("cpu", non_blocking=False) for x in out]

    if sp_size > 1 and os.environ.get("VGPT_SP_WATCHDOG"):
        # debugging aid: after N seconds print this rank's barrier state (epoch, timed-out flag) and the peers' last
        # arrivals, read through a non-blocking side stream so that it works while a kernel of the main stream spins
        def _watch(delay=float(os.environ["VGPT_SP_WATCHDOG"])):
            time.sleep(delay)
            try:
                pg = model._peers
                side = torch.cuda.Stream(device=dev)
                with torch.cuda.stream(side):
                    st = pg._state.to("cpu", non_blocking=True)
                    fl = pg._flags.local[:4 * pg.world].view(torch.int32).to("cpu", non_blocking=True)
                side.synchronize()
                print(f"[watchdog rank {rank}] barrier state (epoch, timed out, ns lo, ns hi) {st.tolist()} peers' arrivals {fl.tolist()}",
                      file=sys.stderr, flush=True)
            except Exception as exc:
                print(f"[watchdog rank {rank}] {type(exc).__name__}: {exc}", file=sys.stderr, flush=True)
        threading.Thread(target=_watch, daemon=True).start()
    for i in range(max(args.warmup, 3)):
        clip_device()
        if os.environ.get("VGPT_SP_TRACE"):
            torch.cuda.synchronize()
            print(f"[rank {rank}] warm-up clip {i} done", file=sys.stderr, flush=True)
    barrier()
    peers = model.sequence_parallel_peers() if sp_size > 1 else None
    barrier_s0 = peers.barrier_seconds() if peers is not None else 0.0
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    s, t = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    s.record()
    for _ in range(args.steps):
        clip_device()
    t.record()
    barrier()
    dt = s.elapsed_time(t) * 1e-3
    clocks = sampler.stop() if rank == 0 else None
    barrier_s = (peers.barrier_seconds() - barrier_s0) if peers is not None else 0.0

    clip_host()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        clip_host()
    barrier()
    dt_e2e = time.perf_counter() - t0

    if world > 1:
        tt = torch.tensor([dt, dt_e2e, barrier_s, -barrier_s], device=dev, dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dt, dt_e2e, barrier_max, barrier_min = float(tt[0]), float(tt[1]), float(tt[2]), -float(tt[3])
    if rank != 0:
        return None
    tokens_per_clip = 2 * n_gen * block * euler * rounds
    videos = world // sp_size
    if vids_rank:
        videos = world * vids_rank
    value = videos * tokens_per_clip * args.steps / dt
    e2e = videos * tokens_per_clip * args.steps / dt_e2e
    step_fl, prefill_fl, clip_fl = algorithmic_flops(dims, n_ctx, n_gen, block, euler)
    e = model.engine()
    # per Euler step: the engine's kernels (the scheduler update rides in the final-layer kernel on one GPU; it is a
    # vgpt_cfg_euler launch of its own, inside the step graph, in sequence-parallel groups and CFG-branch pairs)
    launches = args.steps * rounds * (e.launches_per_prefill + euler * (e.launches_per_predict + (1 if sp_size > 1 else 0)))
    if vids_rank:
        launches *= vids_rank // args.batch

    peaks = {}
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            peaks = json.load(f)
    except Exception:
        pass
    peak_tf = peaks.get("bf16_tflops", 1590.0)
    ncu_traffic = gemm_traffic_from_ncu_summary()
    gemm_rows = e.plan.step.rows
    sec, fl, n_launch = gemm_roofline(model, gemm_rows)
    achieved = fl / sec / 1e12
    roofline = {"bound": "tensor", "kernel": "gemm_bf16_pair_kernel (qkv/o/gate_up/down, M=%d)" % gemm_rows,
                "achieved": achieved, "peak": peak_tf, "unit": "TFLOP/s", "frac": achieved / peak_tf,
                "peak_source": "MEASURED_PEAKS.json bf16_tflops (burst, kernel timed alone)" if peaks else "fallback 1590",
                # DRAM read + write bytes per launch, averaged over the four production launches of a layer (like
                # `achieved`), read from the committed `ncu --set full` summary of the kernels the library runs today
                "traffic": ncu_traffic[0] if gemm_rows == 2064 else None,
                "traffic_note": ncu_traffic[1],
                "avg_launch_us": sec * 1e6, "launches_timed": n_launch,
                "whole_clip_tflops": max(vids_rank, 1) * clip_fl * args.steps / dt / 1e12,
                "whole_clip_frac_of_sustained": max(vids_rank, 1) * clip_fl * args.steps / dt / 1e12 / peaks.get("bf16_tflops_sustained", 1400.0)}

    if args.rollout:          # the per-clip FLOP model assumes one full prefill per clip
        roofline["whole_clip_tflops"] = roofline["whole_clip_frac_of_sustained"] = None

    gpu_eager = None
    if world == 1 and kind == "full" and not args.rollout and not args.no_baselines:
        try:      # the reference path as written, eager bf16 on this GPU (reported next to ours)
            sec_step = gpu_eager_oracle(model, dims, n_ctx, n_gen, H, W, dev)
            gpu_eager = {"value": 2 * n_gen * block / sec_step, "unit": "tokens/s", "ms_per_euler_step": 1e3 * sec_step,
                         "s_per_clip_extrapolated": sec_step * euler,
                         "kind": "oracle port of the reference path, bf16 eager on the same GPU (cuBLAS + SDPA with the "
                                 "dense mask, no KV cache, padded unconditional row); one Euler step timed x3"}
        except Exception as exc:                      # a baseline must never cost the bench line
            gpu_eager = {"error": f"{type(exc).__name__}: {str(exc)[:160]}"}
        torch.cuda.empty_cache()
    threads = os.cpu_count() or 1
    if args.no_baselines:        # exploratory runs (tools/gpu/*.sh): skip the CPU / eager legs, never the default line
        cpu_v, cpu_desc = None, "skipped (--no-baselines)"
    else:
        distinct = 4 if kind == "full" else dims.num_hidden_layers
        cpu_layers = cpu_sample_layers(dims, n_ctx, n_gen, H, W, euler, distinct, threads, 1, 25.0)
        cpu_v, cpu_desc, _ = cpu_reference_sample(dims, n_ctx, n_gen, H, W, euler, distinct, threads, cpu_layers)
        _CPU_SAMPLE_CACHE.clear()
    lat_bytes = 4 * (H // 8) * (W // 8) * 2
    line = {"metric": "next_clip_tokens_per_s", "value": value, "unit": "tokens/s", "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": 1e3 * dt / args.steps,
            "s_per_clip": dt / args.steps, "higher_is_better": True,
            "scaling": "strong" if (sp_size == world and world > 1) or vids_rank else "weak", "vs_baseline": None,
            "dtype": "bf16", "data": "synthetic",
            "config": {"workload": args.config, "model": "Phi-3-mini-class random-init" if kind == "full" else "2 layers / hidden 512",
                       "context_frames": n_ctx, "generated_frames": n_gen, "height": H, "width": W,
                       "euler_steps": euler, "cfg": True, "guidance": GUIDANCE, "prediction_type": "x1",
                       "parallelism": (f"dp{world}: {vids_rank} of {world * vids_rank} videos per rank, {args.batch} videos x 2 CFG "
                                       f"branches per engine pass" if vids_rank else
                                       f"cfg-branch pairs (rank 0 of a pair: conditional sequence, rank 1: unconditional; the "
                                       f"prediction pushed to the peer over NVLink, K/V local) x dp{world // 2}" if cfg_split else
                                       f"sp{sp_size} (rows of one video sharded, K/V pushed to peers over NVLink) x dp{world // sp_size}"
                                       if sp_size > 1 else f"dp{world} (independent videos)"),
                       "rollout": ({"rounds": rounds, "window_frames": n_ctx + n_gen,
                                    "context": "persistent paged K/V cache (each context frame prefilled once)"
                                    if not args.recompute else "recomputed every round (reference flow)"} if args.rollout else None),
                       "l2": "weights 7.2 GB streamed every Euler step (> 126 MB L2); no explicit flush"},
            "clocks": clocks, "gpu_launches": launches,
            "e2e": {"value": e2e, "unit": "tokens/s", "s_per_clip": dt_e2e / args.steps,
                    "h2d_bytes_per_step": max(vids_rank, 1) * (n_ctx + n_gen) * lat_bytes,
                    "d2h_bytes_per_step": max(vids_rank, 1) * n_gen * lat_bytes},
            "roofline": roofline,
            "cpu_baseline": {"value": cpu_v, "unit": "tokens/s", "cores": threads, "kind": "port", "sample": cpu_desc}}
    if gpu_eager is not None:
        line["gpu_eager_baseline"] = gpu_eager
    if sp_size > 1:
        # one video on sp_size GPUs: where the time goes besides the kernels.  Every rank pushes the post-RoPE K / V
        # rows it computes to the other sp_size - 1 ranks (NVLink peer stores fused into rope_kv_append), plus the
        # prediction rows; the flag barriers (one per layer + one per step) are timed on the device (globaltimer).
        rows_step, rows_prefill = 2 * n_gen * block, n_ctx * block
        kv_row = 2 * dims.hidden_size * 2 * dims.num_hidden_layers
        n_steps_timed = args.steps * rounds * euler
        if cfg_split:          # whole sequences per rank: K/V stay local, two barriers per step
            kv_row = 0
        line["sequence_parallel"] = {
            "sp": sp_size, "per_gpu_tflops": clip_fl * args.steps / dt / 1e12 / sp_size,
            "barrier_ms_per_euler_step": {"slowest_rank": 1e3 * barrier_max / n_steps_timed,
                                          "fastest_rank": 1e3 * barrier_min / n_steps_timed,
                                          "note": "device time inside vgpt_peer_barrier kernels (%d per step): waiting for the "
                                                  "slowest peer + NVLink flag round trip; prefill barriers included" % (2 if cfg_split else dims.num_hidden_layers + 1)},
            "nvlink_bytes_per_euler_step": (sp_size - 1) * (rows_step * kv_row + 2 * n_gen * lat_bytes),
            "nvlink_bytes_prefill": (sp_size - 1) * rows_prefill * kv_row}
        roofline["whole_clip_frac_of_sustained"] = roofline["whole_clip_tflops"] / sp_size / peaks.get("bf16_tflops_sustained", 1400.0)
        roofline["whole_clip_frac_note"] = "per GPU: whole-clip TFLOP/s / sp / sustained peak"
    return line


def _free_port():
    import socket
    with socket.socket() as sk:
        sk.bind(("127.0.0.1", 0))
        return sk.getsockname()[1]


def _run_child(cmd, env, timeout_s):
    """Run `cmd` in its own session; on time-out kill the whole process group.  Returns (last JSON line of its stdout,
    None) or (None, reason)."""
    import signal
    p = subprocess.Popen(cmd, env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, start_new_session=True)
    try:
        out, err = p.communicate(timeout=timeout_s)
    except subprocess.TimeoutExpired:
        try:
            os.killpg(p.pid, signal.SIGKILL)
        except ProcessLookupError:
            pass
        p.wait()
        return None, f"timed out after {timeout_s} s (child process group killed)"
    line = None
    for ln in out.splitlines():
        if ln.startswith("{"):
            try:
                line = json.loads(ln)
            except ValueError:
                pass
    if p.returncode != 0 or line is None:
        return None, f"exit {p.returncode}: {(err or out)[-300:]}"
    return line, None


def strong_scaling_child(world, config, timeout_s, parallelism="sp"):
    """One video of `config` over all `world` GPUs -- sequence parallel (rows split, K/V pushed to the peers over NVLink)
    or, with parallelism="cfg" on two GPUs, one CFG branch per rank -- in a CHILD torchrun with a hard time-out, so that
    a stall there can never cost the parent's bench line."""
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), os.path.abspath(__file__), "--gpus", str(world), "--parallelism", parallelism,
           "--config", config, "--steps", "2", "--warmup", "3", "--no-baselines", "--strong", "none"]
    drop = ("RANK", "LOCAL_RANK", "WORLD_SIZE", "LOCAL_WORLD_SIZE", "GROUP_RANK", "GROUP_WORLD_SIZE", "ROLE_RANK", "ROLE_NAME",
            "ROLE_WORLD_SIZE", "MASTER_ADDR", "MASTER_PORT", "OMP_NUM_THREADS")
    env = {k: v for k, v in os.environ.items() if k not in drop and not k.startswith("TORCHELASTIC")}
    t0 = time.perf_counter()
    line, err = _run_child(cmd, env, timeout_s)
    if line is None:
        return {"error": err}
    keep = {k: line.get(k) for k in ("s_per_clip", "value", "unit", "n_gpus", "scaling", "gpu_launches", "sequence_parallel")}
    keep["e2e_s_per_clip"] = line["e2e"]["s_per_clip"]
    keep["parallelism"] = line["config"]["parallelism"]
    keep["child_wall_s"] = time.perf_counter() - t0
    return keep


def strong_scaling_single(configs):
    """The one-GPU denominators of the strong-scaling lines: the same videos on ONE GPU, in-process (fresh model)."""
    from videogpt_b200 import LVMPipeline, LVMProcessor, synth
    from videogpt_b200.synth import SingleIdTagTokenizer as FakeTokenizer
    dev = torch.device("cuda", 0)
    out = {}
    model = build_model(_dims("full"), dev)
    pipe = LVMPipeline(None, model, LVMProcessor(FakeTokenizer()), device=dev)
    for cfg in configs:
        kind, n_ctx, n_gen, H, W, euler = WORKLOADS[cfg]
        lat = [x.to(dev, torch.bfloat16) for x in synth.synthetic_latents(n_ctx + n_gen, H, W, seed=42)]
        kw = dict(num_inference_steps=euler, img_guidance_scale=GUIDANCE, prediction_type="x1")
        try:
            for _ in range(2):
                pipe.next_clip_latents([x.clone() for x in lat[:n_ctx]], n_gen, initial_noise=lat[n_ctx:], **kw)
            torch.cuda.synchronize()
            s, t = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            for _ in range(2):
                pipe.next_clip_latents([x.clone() for x in lat[:n_ctx]], n_gen, initial_noise=lat[n_ctx:], **kw)
            t.record()
            torch.cuda.synchronize()
            sec = s.elapsed_time(t) * 1e-3 / 2
            block = H * W // 256 + 2
            out[cfg] = {"s_per_clip": sec, "value": 2 * n_gen * block * euler / sec, "unit": "tokens/s", "n_gpus": 1,
                        "per_gpu_tflops": algorithmic_flops(_dims(kind), n_ctx, n_gen, block, euler)[2] / sec / 1e12}
        except Exception as exc:
            out[cfg] = {"error": f"{type(exc).__name__}: {str(exc)[:200]}"}
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="cfg2", choices=sorted(WORKLOADS))
    ap.add_argument("--parallelism", default="dp", choices=["dp", "cfg", "sp"],
                    help="N>1: dp = independent videos per rank; cfg = CFG branches split over rank pairs; "
                         "sp = sequence parallel groups of --sp ranks per video")
    ap.add_argument("--sp", type=int, default=0, help="ranks per sequence-parallel group (default: all)")
    ap.add_argument("--batch", type=int, default=4, help="cfg4: videos per engine pass on a rank")
    ap.add_argument("--videos", type=int, default=0, help="cfg4: total videos of the job (default 32 = BASELINE configs[3]; "
                    "a smaller job is a SAMPLE of it, e.g. 8 videos on 1 GPU = the share of two of eight ranks)")
    ap.add_argument("--rollout", type=int, default=0, help="a step = this many clips generated autoregressively in "
                    "latent space (window = the workload's context + clip); 0 = one clip per step")
    ap.add_argument("--recompute", action="store_true", help="--rollout without the persistent K/V cache")
    ap.add_argument("--strong", default="auto", help="strong-scaling block: 'auto' (on for the default cfg2 / dp line: one "
                    "cfg5 and one cfg3 video sharded over all N GPUs, sequence parallel, in a child process with a hard "
                    "time-out; at N = 1 the one-GPU denominators), 'none', or a comma list of configs")
    ap.add_argument("--strong-timeout", type=int, default=240)
    ap.add_argument("--no-baselines", action="store_true", help="skip the cpu_baseline / gpu_eager_baseline legs "
                    "(exploratory runs only; the default line always carries them)")
    args = ap.parse_args()
    if os.environ.get("VGPT_FAULT_DUMP"):      # debugging aid: dump every thread's stack after N seconds and exit
        import faulthandler
        faulthandler.dump_traceback_later(int(os.environ["VGPT_FAULT_DUMP"]), exit=True)
    if args.parallelism == "sp":
        # sequence-parallel ranks spin on each other inside kernels: load every kernel image up front
        # so that no rank stops in the driver (lazy module load) in the middle of a clip
        os.environ.setdefault("CUDA_MODULE_LOADING", "EAGER")
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    line = None
    try:
        line = run_ours(args, rank, world, local_rank)
    finally:
        if world > 1:
            from videogpt_b200 import parallel_states
            parallel_states.destroy_sequence_parallel_group()
            import torch.distributed as dist
            dist.destroy_process_group()
    if line is None:
        return            # ranks > 0 are done; they exit and release their GPUs
    strong = args.strong
    if strong == "auto":
        strong = "cfg5,cfg3" if (args.config == "cfg2" and args.parallelism == "dp" and not args.rollout) else "none"
    if strong != "none":
        # The weak-scaling `value` above is replicas (no collective on the data path).  This block is the part of the
        # multi-GPU design that has one: ONE video sharded over all N GPUs (rows split, K/V rows pushed into the peers'
        # pools over NVLink by the producing kernel, flag barriers), BASELINE configs[2] (cfg3) and configs[4] (cfg5).
        import gc
        gc.collect()
        torch.cuda.empty_cache()
        cfgs = [c for c in strong.split(",") if c in WORKLOADS]
        if world == 1:
            line["strong_scaling"] = {"n_gpus": 1, "mode": "one GPU (denominator of the sp lines at N > 1)", **strong_scaling_single(cfgs)}
        else:
            time.sleep(2.0)      # the other ranks of this job are exiting and releasing their GPUs
            line["strong_scaling"] = {"n_gpus": world, "mode": f"sp{world}: one video over all GPUs, child torchrun, "
                                      f"hard time-out {args.strong_timeout} s",
                                      **{c: strong_scaling_child(world, c, args.strong_timeout) for c in cfgs}}
            if world == 2:       # SURVEY 8(e) axis 1: the two CFG branches of ONE cfg2 video on a rank pair
                line["strong_scaling"]["cfg2_cfg_branch_pair"] = strong_scaling_child(2, "cfg2", args.strong_timeout, "cfg")
    print(json.dumps(line), flush=True)


if __name__ == "__main__":
    main()
