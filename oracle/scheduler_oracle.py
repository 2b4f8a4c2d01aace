"""Oracle restatement of ``LVMScheduler`` (``LVM/scheduler.py:119-208``).

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).
"""
from __future__ import annotations

import copy
from typing import Callable, List, Optional

import torch


def sigma_grid(num_steps: int, time_shifting_factor: float = 1.0, begin_time=None):
    """scheduler.py:120-130: ``linspace`` then ``t / (t + s - s t)``."""
    t = torch.linspace(0 if begin_time is None else begin_time, 1, num_steps + 1)
    return t / (t + time_shifting_factor - time_shifting_factor * t)


def euler_sample(z, func: Callable, model_kwargs: dict, num_steps: int = 50,
                 time_shifting_factor: float = 1.0, prediction_type: str = "v",
                 record: Optional[list] = None):
    """scheduler.py:161-208.  ``z`` is a list of ``[1,C,h,w]`` latents (frame-block path)
    or one batched tensor (``pipeline.__call__`` path).  ``func(z, timesteps, **model_kwargs,
    prediction_type=...)`` returns predictions with the structure of ``z``.
    ``record`` (optional list) receives the per-step velocity actually applied."""
    sigma = sigma_grid(num_steps, time_shifting_factor)
    is_list = isinstance(z, list)
    z = list(z) if is_list else z
    for i in range(num_steps):
        dev = z[0].device
        timesteps = torch.zeros(size=(len(z),)).to(dev) + sigma[i]          # 169
        pred = func(z, timesteps, prediction_type=prediction_type, **model_kwargs)
        s_next, s = sigma[i + 1], sigma[i]
        if prediction_type == "x1":                                            # 180-199
            if is_list:
                pred = [(p - zz) / (1.0 - s) for p, zz in zip(pred, z)]
            else:
                pred = (pred - z) / (1.0 - s)
            if model_kwargs["use_img_cfg"]:
                g = model_kwargs["img_cfg_scale"]
                if is_list:
                    half = len(pred) // 2
                    cond = [u + g * (c - u) for c, u in zip(pred[:half], pred[half:])]
                    pred = cond + cond
                else:
                    c, u = torch.split(pred, len(pred) // 2, dim=0)
                    c = u + g * (c - u)
                    pred = torch.cat([c, c], dim=0)
        if record is not None:
            record.append([p.clone() for p in pred] if is_list else pred.clone())
        if is_list:                                                            # 200-204
            z = [zz + (s_next - s) * p for zz, p in zip(z, pred)]
        else:
            z = z + (s_next - s) * pred
    return z
