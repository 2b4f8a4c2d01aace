"""N > 1 host logic on CPU: world_size-2 (and 4) ``gloo`` process groups on 127.0.0.1."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from videogpt_b200 import parallel


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, fn, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        ret[rank] = fn(rank, world)
    finally:
        dist.destroy_process_group()


def _spawn(world, fn):
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), fn, ret), nprocs=world, join=True)
    return dict(ret)


def _cfg_exchange(rank, world):
    grp = parallel.CfgBranchGroup()
    n_gen = 3
    pred = torch.full((n_gen, 4, 2, 2), float(10 * grp.video_group + grp.branch))
    both = grp.exchange_predictions(pred)
    # layout consumed by vgpt_cfg_euler: cond latents first, then uncond latents
    ok = both.shape == (2 * n_gen, 4, 2, 2) and bool((both[:n_gen] == 10 * grp.video_group).all()) \
        and bool((both[n_gen:] == 10 * grp.video_group + 1).all())
    return ok, grp.video_group, grp.branch, grp.select_branch(["cond", "uncond"])


def test_cfg_branch_exchange_world2():
    out = _spawn(2, _cfg_exchange)
    assert out[0] == (True, 0, 0, "cond") and out[1] == (True, 0, 1, "uncond")


def test_cfg_branch_exchange_world4_two_videos():
    out = _spawn(4, _cfg_exchange)
    assert [out[r][1:3] for r in range(4)] == [(0, 0), (0, 1), (1, 0), (1, 1)] and all(out[r][0] for r in range(4))


def _dp(rank, world):
    mine = parallel.shard_videos(7, rank, world)
    t = parallel.max_over_ranks(1.0 + rank)
    gathered = [None] * world
    dist.all_gather_object(gathered, mine)
    return mine, t, sorted(sum(gathered, []))


def test_video_sharding_and_max_time_world2():
    out = _spawn(2, _dp)
    assert out[0][0] == [0, 2, 4, 6] and out[1][0] == [1, 3, 5]
    assert out[0][1] == out[1][1] == 2.0                 # slowest rank
    assert out[0][2] == list(range(7))                   # every video exactly once


def test_layout_errors():
    with pytest.raises(ValueError):
        parallel.cfg_pair_layout(0, 3)
    with pytest.raises(ValueError):
        parallel.shard_videos(4, 2, 2)
    assert parallel.shard_videos(3, 1, 8) == [1] and parallel.shard_videos(3, 5, 8) == []


# ------------------------------------------------------------------------------------------------
# PeerGroup choreography (sequence parallelism) with the four driver-facing steps replaced by a CPU
# double: handle exchange order, per-rank pointer tables, size-mismatch detection on every rank,
# close() ordering.  The CUDA IPC calls themselves are exercised on 2 GPUs
# (tests/test_sequence_parallel.py).
# ------------------------------------------------------------------------------------------------
def _fake_peer_group(world_ranks, group=None):
    import contextlib
    import struct
    from videogpt_b200 import peer

    class FakePeerGroup(peer.PeerGroup):
        log = []
        _next = 0

        def _device_ctx(self):
            return contextlib.nullcontext()

        def _raw_alloc(self, nbytes):
            FakePeerGroup._next += 1
            return 1000 * (dist.get_rank() + 1) + FakePeerGroup._next        # a "device address" unique per rank

        def _export(self, ptr):
            return struct.pack("<qq", dist.get_rank(), ptr).ljust(64, b"\0")

        def _import(self, handle):
            owner, ptr = struct.unpack("<qq", handle[:16])
            self.log.append(("import", owner, ptr))
            return 10_000_000 + ptr                                          # peer-mapped alias in this process

        def _wrap(self, ptr, nbytes):
            return torch.zeros(nbytes, dtype=torch.uint8)

        def _unmap(self, ptr):
            self.log.append(("unmap", ptr))

        def _free(self, ptr):
            self.log.append(("free", ptr))

        def _sync(self):
            pass

    return FakePeerGroup(world_ranks, group=group, device="cpu")


def _peer_alloc(rank, world):
    grp = _fake_peer_group(list(range(world)))
    buf = grp.alloc(1000)
    mine = 1000 * (rank + 1)
    ok = buf.local.numel() == 1024 and buf.ptrs[rank] // 1000 * 1000 == mine
    for r in range(world):                      # every other rank's buffer appears as its mapped alias
        if r != rank:
            ok &= buf.ptrs[r] >= 10_000_000 and (buf.ptrs[r] - 10_000_000) // 1000 == r + 1
    arr = buf.ptr_array(256)
    ok &= [int(p) for p in arr] == [p + 256 for p in buf.ptrs]
    try:                                        # a size mismatch must raise on EVERY rank, not dead-lock
        grp.alloc(512 if rank == 0 else 2048)
        mismatch = False
    except RuntimeError:
        mismatch = True
    n_log = len(grp.log)
    grp.close()
    closing = [e[0] for e in grp.log[n_log:]]
    return ok, mismatch, closing == sorted(closing, key=lambda k: k != "unmap"), grp.world, grp.rank


def test_peer_group_alloc_exchange_world2():
    out = _spawn(2, _peer_alloc)
    assert all(out[r][:3] == (True, True, True) for r in range(2)) and [out[r][4] for r in range(2)] == [0, 1]


def _peer_subgroups(rank, world):
    from videogpt_b200 import parallel_states as ps
    os.environ["RANK"], os.environ["WORLD_SIZE"] = str(rank), str(world)
    ps.initialize_sequence_parallel_state(2)            # the reference's switch: contiguous groups of 2
    try:
        ranks = dist.get_process_group_ranks(ps.hccl_info.group)
        grp = _fake_peer_group(ranks, group=ps.hccl_info.group)
        buf = grp.alloc(4096)
        partner = ranks[1 - grp.rank]
        return ranks, grp.rank, (buf.ptrs[1 - grp.rank] - 10_000_000) // 1000 == partner + 1
    finally:
        ps.destroy_sequence_parallel_group()


def test_peer_groups_of_two_in_world4():
    out = _spawn(4, _peer_subgroups)
    assert [out[r][0] for r in range(4)] == [[0, 1], [0, 1], [2, 3], [2, 3]]
    assert [out[r][1] for r in range(4)] == [0, 1, 0, 1] and all(out[r][2] for r in range(4))
