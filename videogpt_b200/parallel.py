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
* **sequence** (long contexts): contiguous token chunks per rank like the reference
  (``LVM/model.py:459-464``) with a per-layer K/V all-gather; planned, see DESIGN.md.

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
