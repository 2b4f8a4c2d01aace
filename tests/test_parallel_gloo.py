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
