"""Sequence-parallel process-group state, mirroring ``LVM/acceleration/parallel_states.py``.

``hccl_info`` keeps the reference's field names (``group``, ``world_size``, ``rank``;
parallel_states.py:18-22).  The reference leaves ``world_size`` at 0 until
``initialize_sequence_parallel_state`` runs (and then divides by it: SURVEY.md quirk q1); here
0 means "not initialised" and is treated as a single rank.
"""
from __future__ import annotations

import os

import torch
import torch.distributed as dist


class COMM_INFO:
    def __init__(self):
        self.group = None
        self.world_size = 0
        self.rank = -1
        # host-side rendezvous group of the same ranks over gloo (default; VGPT_SP_HOST_BACKEND=nccl
        # opts out): the peer group's handle exchange and host barriers then launch no GPU kernels at
        # all (an NCCL barrier is an all-reduce kernel that has to share the GPUs with spinning
        # vgpt_peer_barrier kernels, and NCCL's own set-up may block in the driver)
        self.host_group = None
        # what the ranks of the group own (engine.build_plan): "rows" = a chunk of the rows of every sequence, the
        # reference's sequence parallelism (LVM/model.py:459-464); "sequences" = whole sequences -- with two ranks,
        # one CFG branch each (CFG axis): K/V stay local, only the prediction is exchanged
        self.partition = "rows"


hccl_info = COMM_INFO()
_SEQUENCE_PARALLEL_STATE = False


def initialize_sequence_parallel_state(sequence_parallel_size: int, partition: str = "rows"):
    """parallel_states.py:24-37.  ``partition`` (not in the reference): see ``COMM_INFO.partition``."""
    global _SEQUENCE_PARALLEL_STATE
    if partition not in ("rows", "sequences"):
        raise ValueError(f"unknown partition {partition!r}")
    hccl_info.partition = partition
    if sequence_parallel_size > 1:
        _SEQUENCE_PARALLEL_STATE = True
        initialize_sequence_parallel_group(sequence_parallel_size)
    else:
        hccl_info.group, hccl_info.world_size, hccl_info.rank, hccl_info.host_group = None, 1, 0, None


def initialize_cfg_branch_parallel_state():
    """Pairs of ranks share a video, one CFG branch each: rank 0 of a pair runs the conditional sequences (context +
    clip), rank 1 the unconditional ones (clip only, RoPE restarting at 0: quirk q9, so K/V are never shareable)."""
    initialize_sequence_parallel_state(2, partition="sequences")


def get_sequence_parallel_state() -> bool:
    return _SEQUENCE_PARALLEL_STATE


def initialize_sequence_parallel_group(sequence_parallel_size: int):
    """Contiguous groups of ``sequence_parallel_size`` ranks (parallel_states.py:40-54)."""
    rank = int(os.getenv("RANK", "0"))
    world_size = int(os.getenv("WORLD_SIZE", "1"))
    assert world_size % sequence_parallel_size == 0, \
        "world_size must be divisible by sequence_parallel_size"
    hccl_info.world_size = sequence_parallel_size
    hccl_info.rank = rank % sequence_parallel_size
    for i in range(world_size // sequence_parallel_size):
        ranks = list(range(i * sequence_parallel_size, (i + 1) * sequence_parallel_size))
        group = dist.new_group(ranks)
        host_group = dist.new_group(ranks, backend="gloo") if os.getenv("VGPT_SP_HOST_BACKEND", "gloo") == "gloo" else None
        if rank in ranks:
            hccl_info.group, hccl_info.host_group = group, host_group


def destroy_sequence_parallel_group():
    global _SEQUENCE_PARALLEL_STATE
    hccl_info.group, hccl_info.world_size, hccl_info.rank, hccl_info.host_group = None, 0, -1, None
    hccl_info.partition = "rows"
    _SEQUENCE_PARALLEL_STATE = False


def init_env(sequence_parallel_size: int = 1, backend: str = "nccl"):
    """``init_npu_env`` (parallel_states.py:66-81) for NVIDIA: bind the local GPU, join the
    default process group over NCCL (rendezvous from the torchrun environment), build SP groups."""
    local_rank = int(os.getenv("LOCAL_RANK", "0"))
    if torch.cuda.is_available():
        torch.cuda.set_device(local_rank)
    if int(os.getenv("WORLD_SIZE", "1")) > 1 and not dist.is_initialized():
        dist.init_process_group(backend=backend)
    initialize_sequence_parallel_state(sequence_parallel_size)
    return local_rank
