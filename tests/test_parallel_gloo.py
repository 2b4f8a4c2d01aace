"""N > 1 host logic on CPU: world_size-2 (and 4) ``gloo`` process groups on 127.0.0.1."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from videogpt_b200 import parallel, peer as _peer_mod

_REAL_PEER_GROUP = _peer_mod.PeerGroup      # (one test patches peer.PeerGroup with the CPU double)


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

    class FakePeerGroup(_REAL_PEER_GROUP):
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
    # free(): one buffer, collectively -- the peers' mappings go first, then this rank's memory; close() then has
    # nothing left of it
    big = grp.alloc(4096)
    n_log = len(grp.log)
    mapped, mine_big = [p for r, p in enumerate(big.ptrs) if r != rank], big.ptrs[rank]
    grp.free(big)
    freeing = grp.log[n_log:]
    ok &= freeing == [("unmap", p) for p in mapped] + [("free", mine_big)] and big.ptrs == [] and big.local is None
    ok &= mine_big not in grp._owned and not any(p in grp._imported for p in mapped)
    n_log = len(grp.log)
    grp.close()
    closing = [e[0] for e in grp.log[n_log:]]
    ok &= ("free", mine_big) not in grp.log[n_log:]
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


def _cfg_pairs(rank, world):
    from videogpt_b200 import parallel_states as ps
    os.environ["RANK"], os.environ["WORLD_SIZE"] = str(rank), str(world)
    ps.initialize_cfg_branch_parallel_state()           # pairs of ranks, one CFG branch each
    try:
        ranks = dist.get_process_group_ranks(ps.hccl_info.group)
        return ranks, (ranks[0] // 2, ps.hccl_info.rank), ps.hccl_info.partition, ps.hccl_info.world_size
    finally:
        ps.destroy_sequence_parallel_group()
        assert ps.hccl_info.partition == "rows"


def test_cfg_branch_pairs_in_world4():
    """Ranks (2g, 2g+1) serve video group g, branch = rank inside the pair (``parallel.cfg_pair_layout``)."""
    out = _spawn(4, _cfg_pairs)
    assert [out[r][0] for r in range(4)] == [[0, 1], [0, 1], [2, 3], [2, 3]]
    assert [out[r][1] for r in range(4)] == [parallel.cfg_pair_layout(r, 4) for r in range(4)]
    assert all(out[r][2:] == ("sequences", 2) for r in range(4))


# ------------------------------------------------------------------------------------------------
# The whole sequence-parallel HOST flow on CPU: LVM + LVMScheduler under
# initialize_sequence_parallel_state(2), kernels stubbed, peer memory faked, but the host
# rendezvous (PeerGroup.host_barrier = a real gloo barrier) and every decision that changes how
# many collectives a rank enters (plan cache hits, prefill, re-planning) are the real code.
# A rank-inconsistent decision dead-locks here (caught by the timeout) exactly as it would on GPUs.
# ------------------------------------------------------------------------------------------------
def _sp_host_flow(rank, world, partition="rows"):
    from unittest import mock
    from transformers import Phi3Config
    from oracle import processor_oracle as po
    from test_host_dryrun import StubOpsSP
    from videogpt_b200 import LVM, LVMScheduler, engine, model, peer, scheduler, synth
    from videogpt_b200 import parallel_states as ps
    os.environ["RANK"], os.environ["WORLD_SIZE"] = str(rank), str(world)
    stub = StubOpsSP()
    barriers = []
    bf = torch.bfloat16

    def make_group(ranks, group=None, device=None, host_group=None):
        g = _fake_peer_group(ranks, group=group)
        g.barrier = lambda: barriers.append(1)          # the flag-barrier kernel launch
        return g

    def engine_cpu(self):
        if self._engine is None:
            d = self.dims()
            w = engine.EngineWeights(self.state_dict(), d.num_hidden_layers, "cpu")
            self._engine = engine.NextClipEngine(w, d.hidden_size, d.intermediate_size, d.num_hidden_layers,
                                                 d.num_attention_heads, d.rms_norm_eps, d.rope_theta, "cpu",
                                                 self.pos_embed_max_size, self.patch_size, use_cuda_graph=False,
                                                 peers=self.sequence_parallel_peers())
        return self._engine

    ps.initialize_sequence_parallel_state(2, partition=partition)
    try:
        with mock.patch.object(engine, "ops", stub), mock.patch.object(model, "ops", stub), \
                mock.patch.object(scheduler, "ops", stub), mock.patch.object(torch.cuda, "is_available", lambda: True), \
                mock.patch.object(peer, "PeerGroup", make_group), mock.patch.object(model.LVM, "engine", engine_cpu):
            m = LVM(Phi3Config(**synth.REDUCED.phi3_kwargs()), device="cpu", materialize_pos_embed=False).to(bf).eval()
            L = synth.REDUCED.num_hidden_layers
            n_ctx, n_gen, H, W, steps = 3, 2, 64, 96, 2
            d = po.frame_block_inputs(n_ctx, n_gen, H, W, True, 1)
            lat = [x.to(bf) for x in synth.synthetic_latents(n_ctx + n_gen, H, W)]
            per_clip = []
            for clip in range(3):
                # rank 1 churns its allocator between clips: address reuse must not change any decision
                junk = [torch.empty(17 * (clip + 1) * (rank + 1), dtype=bf) for _ in range(5 * rank)]
                mk = dict(input_ids=d["input_ids"].clone(), input_img_latents=[x.clone() for x in lat[:n_ctx]],
                          input_image_sizes=d["input_image_sizes"], attention_mask=None,
                          position_ids=d["position_ids"].clone(), denoise_image_sizes=d["denoise_image_sizes"],
                          time_emb_inx=d["time_emb_inx"], img_cfg_scale=1.5, use_img_cfg=True, use_kv_cache=False,
                          offload_model=False, vae=None)
                before = len(barriers)
                LVMScheduler(steps)([x.clone() for x in lat[n_ctx:]] * 2, m.frame_block_forward_with_cfg, mk,
                                    prediction_type="x1")
                per_clip.append(len(barriers) - before)
                del junk, mk
            e = m._engine
            return per_clip, e.plan.shard, (e.plan.prefix.rows, e.plan.step.rows), L, e.plan.partition
    finally:
        ps.destroy_sequence_parallel_group()


@pytest.mark.timeout(240)
def test_sequence_parallel_host_flow_world2():
    out = _spawn(2, _sp_host_flow)
    L = out[0][3]
    # every clip: prefill (L barriers) + 2 steps x (L + 1) barriers, on BOTH ranks, every time
    assert out[0][0] == out[1][0] == [L + 2 * (L + 1)] * 3
    assert out[0][1] == (0, 2) and out[1][1] == (1, 2)
    assert out[0][2][0] + out[1][2][0] == 3 * 26 and out[0][2][1] + out[1][2][1] == 2 * 2 * 26


def _cfg_pair_host_flow(rank, world):
    return _sp_host_flow(rank, world, partition="sequences")


@pytest.mark.timeout(240)
def test_cfg_branch_pair_host_flow_world2():
    """Same flow with one CFG branch per rank (``partition="sequences"``): rank 1 has no context rows, yet both ranks
    enter the same host rendezvous and the same two barriers per Euler step, every clip."""
    out = _spawn(2, _cfg_pair_host_flow)
    assert out[0][0] == out[1][0] == [2 * 2] * 3
    assert out[0][1] == (0, 2) and out[1][1] == (1, 2) and out[0][4] == out[1][4] == "sequences"
    assert out[0][2] == (3 * 26, 2 * 26) and out[1][2] == (0, 2 * 26)
