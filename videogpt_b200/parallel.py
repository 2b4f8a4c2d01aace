"""Multi-GPU partitioning of the next-clip path (one process per GPU, ``torch.distributed``).

The path shards along three independent axes; none of them puts a collective
library call on the transformer data path:

* **videos** (``shard_videos``): independent clips are dealt round-robin to ranks -- pure data
  parallelism, the only communication is gathering results (or nothing at all).
* **CFG branches** (``parallel_states.initialize_cfg_branch_parallel_state()``; ``cfg_pair_layout``
  names the pairs): rank 0 of a pair runs the conditional sequence (context + generated clip), rank 1
  the unconditional one (generated clip only, RoPE restarting at 0: quirk q9, so K/V are never
  shareable).  Per Euler step the two ranks exchange ONE tensor -- the raw prediction,
  ``n_gen x 4 x h/8 x w/8`` bf16 (32 KB at 256x256): the final-layer kernel of each rank stores its
  half into both ranks' buffers over NVLink, a flag barrier orders the stores, and both ranks apply
  the same x1->v / CFG / Euler update inside the step graph, so their latents stay bit-identical.
* **sequence** (long contexts / one video on several GPUs,
  ``parallel_states.initialize_sequence_parallel_state(P)``): rows of every sequence are dealt to
  the ranks of ``hccl_info.group`` (the reference's switch, ``LVM/model.py:459-464``); the K/V
  all-gather is fused into the producing kernels as NVLink peer stores.

The last two are the same mechanism (``peer.py``, ``csrc/peer.cu``, ``engine.build_plan(shard=,
partition=)``) with different ownership -- whole sequences or chunks of rows -- and live in the
engine / model, not here: see DESIGN.md section 7.

The reference's own multi-GPU inference is DeepSpeed-Ulysses all-to-all (4 collectives per
layer, ``LVM/transform/sdpa_transform.py:126-156``).
"""
from __future__ import annotations

from typing import List

import torch
import torch.distributed as dist


def shard_videos(n_videos: int, rank: int, world: int) -> List[int]:
    """Indices of the videos rank ``rank`` of ``world`` processes (round-robin: balanced to +-1)."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError(f"bad rank/world {rank}/{world}")
    return list(range(rank, n_videos, world))


def cfg_pair_layout(rank: int, world: int):
    """(video_group, branch) of a rank when CFG branches are split over rank pairs:
    ranks (2g, 2g+1) serve video group g; branch 0 = conditional, 1 = unconditional -- the contiguous
    groups of two that ``initialize_cfg_branch_parallel_state`` builds, and ``engine.sequence_owner``
    of a two-sequence batch."""
    if world % 2 != 0:
        raise ValueError("CFG-branch parallelism needs an even number of ranks")
    return rank // 2, rank % 2


def max_over_ranks(seconds: float, device=None) -> float:
    """Wall/device time reported for a multi-GPU run = the slowest rank."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return seconds
    t = torch.tensor([seconds], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t[0])
