"""``LVMScheduler``: drop-in for the reference's flow-matching Euler sampler
(``LVM/scheduler.py:119-208``): same constructor, sigma grid and call signature.

When ``func`` is the ``frame_block_forward_with_cfg`` of a ``videogpt_b200.LVM`` the whole
loop runs on the engine: latents stay in one device buffer, each step is one CUDA-graph replay
of the denoising forward plus the fused x1->v / CFG / Euler kernel, and nothing syncs with the
host.  Any other ``func`` (seam S2 of SURVEY.md 8(b)) goes through the generic loop, which still
applies the update with the CUDA kernel.  CPU tensors are rejected: there is no CPU path.
"""
from __future__ import annotations

from typing import Callable, Optional

import torch

from . import ops


class LVMScheduler:
    def __init__(self, num_steps: int = 50, time_shifting_factor: int = 1, begin_time=None):
        self.num_steps = num_steps
        self.time_shift = time_shifting_factor
        t = torch.linspace(0 if begin_time is None else begin_time, 1, num_steps + 1)
        self.sigma = t / (t + time_shifting_factor - time_shifting_factor * t)
        self.record_velocity: Optional[list] = None     # tests: per-step velocity of the cond half

    # step scalars exactly as the reference forms them (fp32 tensor arithmetic, scheduler.py:178-204)
    def _scalars(self, i: int):
        sigma, sigma_next = self.sigma[i], self.sigma[i + 1]
        return float(1.0 - sigma), float(sigma_next - sigma)

    def __call__(self, z, func: Callable, model_kwargs: dict, use_kv_cache: bool = True,
                 offload_kv_cache: bool = True, prediction_type: str = "v", vae=None, noise_level=None):
        if noise_level is not None:
            z = z * noise_level + torch.randn_like(z) * (1 - noise_level)
        model = getattr(func, "__self__", None)
        from .model import LVM
        name = getattr(func, "__name__", "")
        if isinstance(model, LVM) and name == "frame_block_forward_with_cfg" and isinstance(z, list):
            out = self._run_engine(z, model, model_kwargs, prediction_type)
        elif isinstance(model, LVM) and name == "forward_with_cfg" and torch.is_tensor(z):
            out = self._run_engine(z, model, model_kwargs, prediction_type)
        else:
            out = self._run_generic(z, func, model_kwargs, prediction_type)
        # (The reference ends with `del cache; torch.cuda.empty_cache(); gc.collect()` -- scheduler.py:205-207 -- to
        # drop the throw-away DynamicCache of quirk q14.  There is no such garbage here, and a full collection costs
        # ~90 ms of host time in a process that holds torch + transformers: hidden behind queued GPU work on one GPU,
        # but fully exposed in a peer group, whose ranks synchronise with their GPUs at every clip boundary --
        # 0.599 instead of 0.53 s/clip for a CFG-branch pair at cfg2, tools/gpu/r02s.sh.)
        return out

    # ---- engine loop -----------------------------------------------------------------------------
    def _run_engine(self, z, model, mk: dict, prediction_type: str):
        is_list = isinstance(z, list)
        lat_h, lat_w = z[0].shape[-2:]
        if is_list:
            e = model.prepare_frame_block(mk["input_ids"], mk["input_img_latents"], mk["input_image_sizes"],
                                          mk["attention_mask"], mk["position_ids"], mk["denoise_image_sizes"],
                                          mk["time_emb_inx"], lat_h, lat_w)
        else:
            e = model.prepare_single_frame(mk["input_ids"], mk["input_img_latents"], mk["input_image_sizes"],
                                           mk["attention_mask"], mk["position_ids"], lat_h, lat_w)
        return self.run_prepared(e, z, bool(mk["use_img_cfg"]), float(mk["img_cfg_scale"]), prediction_type)

    def run_prepared(self, e, z, use_cfg: bool, guidance: float, prediction_type: str):
        """The sampling loop on an engine whose plan is set and whose prefix is prefilled
        (``_run_engine`` after ``LVM.prepare_*``; ``rollout.LatentRollout`` after its own plan)."""
        is_list = isinstance(z, list)
        lat_h, lat_w = z[0].shape[-2:]
        n = e.plan.n_latents
        assert len(z) == n
        e.z.copy_(torch.cat([t.reshape(1, 4, lat_h, lat_w) for t in z], 0) if is_list else z)
        x1 = prediction_type == "x1"
        n_half = n // 2 if use_cfg else n
        # per-step inputs of the whole clip in one table: [sigma_i x n | 1 - sigma_i, sigma_{i+1} - sigma_i, guidance]
        # (the scalars exactly as the reference forms them); one device-to-device copy per step selects row i
        rows = [[float(self.sigma[i])] * n + list(self._scalars(i)) + [guidance] for i in range(self.num_steps)]
        table = torch.tensor(rows, dtype=torch.float32).to(e.step_inputs.device, non_blocking=True)
        e.uniform_t = True               # one sigma for every latent (scheduler.py:171)
        # The update is part of the step graph: inside the final-layer kernel on one GPU; in a peer group (sequence
        # parallel, CFG-branch pairs) as vgpt_cfg_euler behind the barrier that orders the peers' prediction stores,
        # applied by every rank to its full copy of z.
        e.euler_mode = (use_cfg, x1)
        try:
            for i in range(self.num_steps):
                e.step_inputs.copy_(table[i])
                e.predict()
                if self.record_velocity is not None:
                    self.record_velocity.append(e.vel[:n_half].clone())
        finally:
            e.uniform_t = False
            e.euler_mode = None
        if e.peers is not None:
            # peer group: a barrier that timed out (a peer died or fell behind by ~10 s) lets this rank run
            # on K/V and prediction rows its peers never delivered -- never return such latents as a result
            e.peers.check()
        if not is_list:
            return e.z.clone()
        out = [e.z[i:i + 1].clone() for i in range(n)]
        for i in range(n):
            z[i] = out[i]
        return out

    # ---- generic loop (any func with the reference's callback signature) --------------------------
    def _run_generic(self, z, func, mk: dict, prediction_type: str):
        is_list = isinstance(z, list)
        zs = torch.cat(z, 0) if is_list else z
        if not zs.is_cuda:
            raise RuntimeError("videogpt_b200.LVMScheduler needs CUDA latents (no CPU fallback)")
        zs = zs.to(torch.bfloat16).contiguous().clone()
        use_cfg = bool(mk.get("use_img_cfg", False))
        for i in range(self.num_steps):
            cur = [zs[j:j + 1] for j in range(zs.shape[0])] if is_list else zs
            timesteps = torch.zeros(size=(len(cur),), device=zs.device) + self.sigma[i]
            pred, _cache = func(cur, timesteps, past_key_values=None, prediction_type=prediction_type, **mk)
            pred = (torch.cat(pred, 0) if is_list else pred).to(torch.bfloat16).contiguous()
            oms, ds = self._scalars(i)
            if prediction_type == "x1":
                ops.cfg_euler(zs, pred, use_cfg, True, oms, ds, float(mk.get("img_cfg_scale", 1.0)))
            else:   # v mode: the model already combined the branches (model.py:554-562)
                ops.cfg_euler(zs, pred, False, False, oms, ds, 1.0)
        if is_list:
            out = [zs[j:j + 1].clone() for j in range(zs.shape[0])]
            for j in range(len(out)):
                z[j] = out[j]
            return out
        return zs
