"""Multi-GPU partitioning of the next-clip path (one process per GPU, ``torch.distributed``).

The path shards along three independent axes; none of them needs a collective
on the transformer data path:

* **videos** (``shard_videos``): independent clips are dealt round-robin to ranks -- pure data
  parallelism, the only communication is gathering results (or nothing at all).
* **CFG branches** (``CfgBranchGroup``): rank 0 of a pair runs the conditional sequence (context
  + generated clip), rank 1 the unconditional one (generated clip only, RoPE restarting at 0:
  quirk q9, so K/V are never shareable).  Per Euler step the two ranks exchange ONE tensor -- the
  raw prediction, ``n_gen x 4 x h/8 x w/8`` bf16 (32 KB at 256x256) -- with an all-gather, then
  both apply the same x1->v / CFG / Euler update, so their latents stay bit-identical.
* **sequence** (long contexts / one video on several GPUs): rows of every sequence are dealt to
  the ranks of ``hccl_info.group`` (the reference's switch, ``LVM/model.py:459-464``); the K/V
  all-gather is fused into the producing kernels as NVLink peer stores (``peer.py``,
  ``csrc/peer.cu``).  It lives in the engine / model, not here: see DESIGN.md section 7.

The reference's own multi-GPU inference is DeepSpeed-Ulysses all-to-all (4 collectives per
layer, ``LVM/transform/sdpa_transform.py:126-156``).
"""
from __future__ import annotations

from typing import List, Optional, Sequence

import torch
import torch.distributed as dist


def shard_videos(n_videos: int, rank: int, world: int) -> List[int]:
    """Indices of the videos rank ``rank`` of ``world`` processes (round-robin: balanced to +-1)."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError(f"bad rank/world {rank}/{world}")
    return list(range(rank, n_videos, world))


def cfg_pair_layout(rank: int, world: int):
    """(video_group, branch) of a rank when CFG branches are split over rank pairs:
    ranks (2g, 2g+1) serve video group g; branch 0 = conditional, 1 = unconditional."""
    if world % 2 != 0:
        raise ValueError("CFG-branch parallelism needs an even number of ranks")
    return rank // 2, rank % 2


class CfgBranchGroup:
    """Process group of the two ranks that share one video's CFG branches."""

    def __init__(self, rank: Optional[int] = None, world: Optional[int] = None):
        self.rank = dist.get_rank() if rank is None else rank
        self.world = dist.get_world_size() if world is None else world
        self.video_group, self.branch = cfg_pair_layout(self.rank, self.world)
        self.group = None
        for g in range(self.world // 2):           # every rank must create every group
            grp = dist.new_group([2 * g, 2 * g + 1])
            if g == self.video_group:
                self.group = grp

    def select_branch(self, per_branch: Sequence):
        """This rank's element of a ``[cond, uncond]`` pair (sequence spec, latents, ...)."""
        assert len(per_branch) == 2
        return per_branch[self.branch]

    def exchange_predictions(self, pred_local: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """All-gather the branch predictions into the ``[cond latents | uncond latents]`` layout the
        CFG/Euler kernel (``vgpt_cfg_euler``) consumes.  ``pred_local``: ``[n_gen, 4, h, w]``."""
        if out is None:
            out = pred_local.new_empty((2 * pred_local.shape[0],) + tuple(pred_local.shape[1:]))
        dist.all_gather_into_tensor(out, pred_local.contiguous(), group=self.group)
        return out


def max_over_ranks(seconds: float, device=None) -> float:
    """Wall/device time reported for a multi-GPU run = the slowest rank."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return seconds
    t = torch.tensor([seconds], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t[0])


# ------------------------------------------------------------------------------------------------
# CFG-branch parallel sampling: two ranks per video, one 32 KB all-gather per Euler step
# ------------------------------------------------------------------------------------------------
def branch_spec(specs, branch: int):
    """Sequence spec of one CFG branch with latent / context numbering local to its rank.
    ``specs`` = [cond, uncond] from ``engine.frame_block_specs``."""
    import copy
    import numpy as np
    from . import ops
    sp = copy.deepcopy(specs[branch])
    lat_ids = sorted(l for l, _ in sp.latent_rows)
    base = lat_ids[0]
    uses_latent = (sp.kinds == ops.ROW_TIME) | (sp.kinds == ops.ROW_NOISY_PATCH)
    sp.arg_a = np.where(uses_latent, sp.arg_a - base, sp.arg_a).astype(np.int32)
    sp.latent_rows = [(l - base, r) for l, r in sp.latent_rows]
    n_ctx = int((sp.kinds == ops.ROW_CONTEXT_PATCH).any()) and int(sp.arg_a[sp.kinds == ops.ROW_CONTEXT_PATCH].max()) + 1
    if n_ctx:
        ctx0 = int(sp.arg_a[sp.kinds == ops.ROW_CONTEXT_PATCH].min())
        sp.arg_a = np.where(sp.kinds == ops.ROW_CONTEXT_PATCH, sp.arg_a - ctx0, sp.arg_a).astype(np.int32)
        n_ctx -= ctx0
    return sp, len(lat_ids), n_ctx


@torch.no_grad()
def sample_cfg_split(model, scheduler, z: List[torch.Tensor], model_kwargs: dict, grp: CfgBranchGroup,
                     prediction_type: str = "x1") -> List[torch.Tensor]:
    """``LVMScheduler`` loop with the conditional and unconditional branch on different GPUs.

    ``z`` and ``model_kwargs`` are the same objects the single-GPU path takes (both rows present on
    both ranks; the host work is replicated like in the reference's SP ranks).
    Returns the ``n_gen`` generated latents (identical on both ranks)."""
    from . import engine as eng, ops
    mk = model_kwargs
    assert mk["use_img_cfg"], "CFG-branch parallelism needs guidance on"
    lat_h, lat_w = z[0].shape[-2:]
    specs, n_lat, n_ctx_total = eng.frame_block_specs(mk["input_ids"], mk["position_ids"], mk["input_image_sizes"],
                                                     mk["denoise_image_sizes"], mk["time_emb_inx"])
    assert len(specs) == 2 and n_lat % 2 == 0
    n_gen = n_lat // 2
    sp, n_local, n_ctx = branch_spec(specs, grp.branch)
    e = model.engine()
    layout = ("cfg-split", grp.branch, sp.codes.tobytes(), sp.positions.tobytes(), sp.kinds.tobytes(),
              sp.arg_a.tobytes(), lat_h, lat_w)
    if model._layout_key != layout or e.plan is None:    # same geometry as the last clip: keep plan + graph
        e.set_plan(eng.build_plan([sp], n_local, n_ctx, lat_h, lat_w, e.device))
        model._layout_key = layout
    model._plan_key = None                               # the engine does not hold a 2-row plan
    ctx = torch.cat([x.reshape(1, 4, lat_h, lat_w) for x in mk["input_img_latents"]], 0) if n_ctx else None
    e.prefill(ctx)
    z_all = torch.cat([t.reshape(1, 4, lat_h, lat_w) for t in z], 0).to(e.device, eng.ACT_DTYPE).contiguous()
    pred_all = torch.empty_like(z_all)
    for i in range(scheduler.num_steps):
        e.z.copy_(z_all[:n_gen])                         # both halves of z_all are identical (quirk q7)
        e.t.fill_(float(scheduler.sigma[i]))
        e.predict()
        grp.exchange_predictions(e.pred, pred_all)
        oms, ds = scheduler._scalars(i)
        ops.cfg_euler(z_all, pred_all, True, prediction_type == "x1", oms, ds, float(mk["img_cfg_scale"]))
    return [z_all[i:i + 1].clone() for i in range(n_gen)]
