"""Latent-space autoregressive rollout with a persistent paged K/V cache.

The reference's rollout (``LVM/pipeline.py:418-424, 485-500``) decodes every generated clip to
uint8 PIL images, re-encodes ALL frames of the window through the VAE and recomputes the K/V of
every context frame, every round.  Here latents are carried forward and **every context frame
is processed exactly once**, in the round in which it first becomes context: its K/V stay in the
paged pool, later rounds only prefill the frames generated in the round before
(``SequenceSpec.n_cached``), and the page table is a sliding window over absolute pages --
pages that fall wholly behind the window go back to the free list.

What it computes (its parity definition, pinned against the oracle by
``tests/test_emu_host_numerics.py`` on CPU and ``tests/test_zz_rollout_gpu.py`` on the GPU):

* While the window still holds the whole history (``frames + gen_num <= max_frame_window``)
  a round is EXACTLY the reference's round on the same context latents: context frames are
  frame-causal, so a frame's K/V do not depend on frames that come later, and RoPE positions are
  the same sequential token indices.
* Once frames are evicted, the reference would restart positions at 0 and recompute the kept
  frames without the dropped ones.  This rollout instead keeps what it cached: the round is the
  reference's forward over the FULL history with absolute positions and a mask in which (a)
  the clip being generated sees the last ``max_frame_window - gen_num`` context frames, and (b)
  every context frame keeps the view it had in the round that cached it.  Attention scores only
  depend on position differences, so nothing inside the window moves relative to anything else.
* ``clean_image_noise_level`` is applied once, when a generated clip becomes context (the
  reference re-noises every re-encoded frame every round).

Under sequence parallelism (``initialize_sequence_parallel_state(P)``, BASELINE configs[2]) the
rounds run on row-sharded plans: each rank prefills its chunk of the NEW context rows and stores
their K/V into every peer's pool, which persists across rounds like the single-GPU one; latents
are replicated, so every rank carries the same history (``clean_image_noise_level > 0`` would
need a shared generator and is refused).
"""
from __future__ import annotations

from typing import List, Optional

import numpy as np
import torch

from . import engine as eng
from . import ops
from .ops import PAGE_TOKENS
from .scheduler import LVMScheduler


def _tag_id(tokenizer, text: str) -> int:
    ids = list(tokenizer(text).input_ids)
    if ids and ids[0] == 1:                       # BOS strip, LVM/processor.py:140-142
        ids = ids[1:]
    assert len(ids) == 1, f"{text!r} must tokenise to exactly one id (LVM/processor.py:138-142)"
    return int(ids[0])


def window_start(n_hist: int, gen_num: int, max_frame_window: int) -> int:
    """First context frame the clip generated after ``n_hist`` frames can see
    (``input_images[gen_num + len - max_frame_window:]``, LVM/pipeline.py:421-422)."""
    return max(0, n_hist + gen_num - max_frame_window)


class LatentRollout:
    def __init__(self, model, processor, gen_num: int, max_frame_window: int = 16, num_inference_steps: int = 50,
                 img_guidance_scale: float = 1.6, use_img_guidance: bool = True, time_shifting_factor: float = 1.0,
                 prediction_type: str = "v", clean_image_noise_level: float = 0.0, rounds_hint: int = 16):
        if max_frame_window <= gen_num:
            raise ValueError("max_frame_window must leave room for at least one context frame")
        self.model, self.gen, self.window = model, int(gen_num), int(max_frame_window)
        self.steps, self.tsf, self.pt = num_inference_steps, time_shifting_factor, prediction_type
        self.guidance = float(img_guidance_scale)
        self.use_cfg = bool(use_img_guidance) and img_guidance_scale != 1
        self.noise_level = float(clean_image_noise_level or 0.0)
        self.rounds_hint = rounds_hint
        tok = processor.text_tokenizer
        self.id_open, self.id_close, self.id_diff = (_tag_id(tok, t) for t in ("<img>", "</img>", "<|diffusion|>"))
        self.n_hist = 0            # frames of history (all of them context for the next clip)
        self.n_done = 0            # frames whose K/V are in the pool
        self.pending: List[torch.Tensor] = []     # latents of frames [n_done, n_hist)
        self.engine = None
        self._phys = {}            # absolute page -> physical page (conditional sequence)
        self._free: List[int] = []
        self.prefilled_frames = 0  # statistics: context frames pushed through the transformer so far
        self._plan = self._kv = None

    # ---- set-up ----------------------------------------------------------------------------------
    def start(self, context_latents: List[torch.Tensor]):
        if not context_latents:
            raise ValueError("a rollout needs context frames")
        self.lat_h, self.lat_w = (int(x) for x in context_latents[0].shape[-2:])
        self.bl = (self.lat_h // 2) * (self.lat_w // 2) + 2          # tokens per frame block
        e = self.engine = self.model.engine()
        self.model.invalidate_plan_cache()                            # the engine's plan is ours now
        self.dev = e.device
        self.pending = [x.to(self.dev, eng.ACT_DTYPE).reshape(1, 4, self.lat_h, self.lat_w) for x in context_latents]
        self.n_hist, self.n_done = len(self.pending), 0
        # capacities that never change, so that later rounds refresh the plan in place
        live_tokens = max(self.window, len(self.pending) + self.gen) * self.bl
        self.max_pages = (live_tokens + PAGE_TOKENS - 1) // PAGE_TOKENS + 1
        self.uncond_pages = (self.gen * self.bl + PAGE_TOKENS - 1) // PAGE_TOKENS if self.use_cfg else 0
        self.pool_pages = self.max_pages + self.uncond_pages
        self._free = list(range(self.uncond_pages, self.pool_pages))
        self._phys = {}
        e.rope_reserve = max(e.rope_reserve, (len(self.pending) + (self.rounds_hint + 1) * self.gen) * self.bl)
        self.round = 0
        return self

    # ---- one round -------------------------------------------------------------------------------
    def _specs(self):
        bl, gen, n_hist = self.bl, self.gen, self.n_hist
        ws = window_start(n_hist, gen, self.window)
        c = max(self.n_done, ws)                     # first frame to prefill this round
        p0 = (ws * bl) // PAGE_TOKENS                # first live absolute page
        a0 = p0 * PAGE_TOKENS
        end = (n_hist + gen) * bl
        a = np.arange(a0, end)
        f, o = a // bl, a % bl
        ctx = f < n_hist
        r_ctx = np.where(o == 0, 0, np.where(o == bl - 1, 2, 1))
        codes = np.where(f < ws, eng.INT_MAX, np.where(ctx, 4 * (f - ws) + r_ctx, 4 * (n_hist - ws) + np.minimum(o, 2)))
        kinds = np.full(len(a), ops.ROW_TOKEN, np.int32)
        arg_a = np.zeros(len(a), np.int32)
        arg_b = np.zeros(len(a), np.int32)
        new = ctx & (f >= c)
        arg_a[new & (o == 0)] = self.id_open
        arg_a[new & (o == bl - 1)] = self.id_close
        img = new & (o > 0) & (o < bl - 1)
        kinds[img] = ops.ROW_CONTEXT_PATCH
        arg_a[img] = (f - c)[img]
        arg_b[img] = (o - 1)[img]

        def gen_rows(kinds, arg_a, arg_b, sel, j, o, lat0):
            arg_a[sel & (o == 0)] = self.id_diff
            t = sel & (o == 1)
            kinds[t] = ops.ROW_TIME
            arg_a[t] = (lat0 + j)[t]
            px = sel & (o >= 2)
            kinds[px] = ops.ROW_NOISY_PATCH
            arg_a[px] = (lat0 + j)[px]
            arg_b[px] = (o - 2)[px]

        gen_rows(kinds, arg_a, arg_b, ~ctx, f - n_hist, o, 0)
        n_prefix = n_hist * bl - a0
        # physical pages: keep what is mapped, map what is new, release what fell behind the window
        last = (end - 1) // PAGE_TOKENS
        for page in [p for p in self._phys if p < p0]:
            self._free.append(self._phys.pop(page))
        for page in range(p0, last + 1):
            if page not in self._phys:
                if not self._free:
                    raise RuntimeError("rollout: K/V pool exhausted (window arithmetic is off)")
                self._phys[page] = self._free.pop(0)
        pages = np.array([self._phys[p] for p in range(p0, last + 1)], np.int32)
        cond = eng.SequenceSpec(n_prefix=n_prefix, n_active=gen * bl, positions=a.astype(np.int32),
                                codes=codes.astype(np.int32), kinds=kinds, arg_a=arg_a, arg_b=arg_b,
                                latent_rows=[(j, n_prefix + j * bl + 2) for j in range(gen)],
                                n_cached=c * bl - a0, pages=pages)
        specs = [cond]
        if self.use_cfg:                             # unconditional row: the clip alone, positions from 0 (quirk q9)
            u = np.arange(gen * bl)
            fu, ou = u // bl, u % bl
            k2 = np.full(len(u), ops.ROW_TOKEN, np.int32)
            a2 = np.zeros(len(u), np.int32)
            b2 = np.zeros(len(u), np.int32)
            gen_rows(k2, a2, b2, np.ones(len(u), bool), fu, ou, gen)
            specs.append(eng.SequenceSpec(n_prefix=0, n_active=gen * bl, positions=u.astype(np.int32),
                                          codes=np.minimum(ou, 2).astype(np.int32), kinds=k2, arg_a=a2, arg_b=b2,
                                          latent_rows=[(gen + j, j * bl + 2) for j in range(gen)],
                                          pages=np.arange(self.uncond_pages, dtype=np.int32)))
        return specs, c

    @torch.no_grad()
    def next_clip(self, seed: Optional[int] = None, initial_noise: Optional[List[torch.Tensor]] = None):
        """Generate the next ``gen_num`` latents ``[1,4,h,w]`` and append them to the history."""
        e, gen = self.engine, self.gen
        if e is None:
            raise RuntimeError("call start(context_latents) first")
        if self.model.engine() is not e:
            raise RuntimeError("the model rebuilt its engine (weights moved or changed): start() the rollout again")
        if self.round > 0 and (e.plan is not self._plan or e.kv.data_ptr() != self._kv):
            raise RuntimeError("the engine was used for something else since the last round: its K/V pool no longer "
                               "holds this rollout's context; start() again")
        specs, c = self._specs()
        new_ctx = self.pending[c - self.n_done:]
        shard = None if e.peers is None else (e.peers.rank, e.peers.world)     # sequence parallel: rows dealt to the ranks
        if shard is not None and seed is None and initial_noise is None:
            raise ValueError("under sequence parallelism every rank must draw the same noise: pass seed or initial_noise")
        plan = eng.build_plan(specs, gen * (2 if self.use_cfg else 1), len(new_ctx), self.lat_h, self.lat_w, self.dev,
                              shard=shard, max_pages=self.max_pages, pool_pages=self.pool_pages)
        if plan.max_pos > e.rope_reserve:
            e.rope_reserve = 2 * plan.max_pos
        e.set_plan(plan, keep_kv=self.round > 0)
        self.model.invalidate_plan_cache()
        e.prefill(torch.cat(new_ctx, 0) if new_ctx else None)
        self.prefilled_frames += len(new_ctx)
        if initial_noise is not None:
            z = [x.to(self.dev, eng.ACT_DTYPE).reshape(1, 4, self.lat_h, self.lat_w) for x in initial_noise]
        else:
            g = torch.Generator(device=self.dev).manual_seed(seed) if seed is not None else None
            z = [torch.randn(1, 4, self.lat_h, self.lat_w, device=self.dev, generator=g).to(eng.ACT_DTYPE)
                 for _ in range(gen)]
        assert len(z) == gen
        z = z * (2 if self.use_cfg else 1)
        sch = LVMScheduler(num_steps=self.steps, time_shifting_factor=self.tsf)
        out = sch.run_prepared(e, z, self.use_cfg, self.guidance, self.pt)[:gen]
        # the clip becomes context of the next round
        a = self.noise_level
        if a and e.peers is not None:
            raise NotImplementedError("clean_image_noise_level > 0 under sequence parallelism (ranks would draw different noise)")
        self.pending = [((1 - a) * x + a * torch.randn_like(x)) if a else x for x in out]
        self.n_done, self.n_hist = self.n_hist, self.n_hist + gen
        self._plan, self._kv = e.plan, e.kv.data_ptr()
        self.round += 1
        return out
