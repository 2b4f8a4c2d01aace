"""Sequence parallelism (sequence axis; reference: LVM/model.py:459-474 chunking +
LVM/transform/sdpa_transform.py:126-156 Ulysses all-to-alls).

Here a rank owns a contiguous chunk of the rows of every sequence and stores its K/V rows and its
predictions into every peer (videogpt_b200/peer.py); every row-wise op is independent of the
partition, so a sharded run must reproduce the single-GPU run BIT FOR BIT:

* CPU: the union of the ranks' row arrays is exactly the unsharded plan (any world size);
* 1 GPU: ``world`` virtual ranks (LocalPeerGroup, lockstep) == the unsharded engine, bit-exact,
  through the real kernels (vgpt_rope_kv_append_peers, vgpt_final_layer_rows);
* 2 GPUs: LVM + LVMScheduler under ``initialize_sequence_parallel_state(2)`` (real IPC peer
  memory + the barrier kernel, CUDA graphs on) == the single-GPU run, bit-exact.
"""
import os
import socket

import numpy as np
import pytest
import torch

from oracle import processor_oracle as po
from videogpt_b200 import engine as eng, synth


def _specs(n_ctx, n_gen, H, W):
    d = po.frame_block_inputs(n_ctx, n_gen, H, W, True, 1)
    specs, n_lat, n_ctx_lat = eng.frame_block_specs(d["input_ids"], d["position_ids"], d["input_image_sizes"],
                                                   d["denoise_image_sizes"], d["time_emb_inx"])
    return d, specs, n_lat, n_ctx_lat


@pytest.mark.parametrize("world", [2, 3, 8])
@pytest.mark.parametrize("geom", [(4, 4, 256, 256), (3, 2, 64, 96), (1, 1, 64, 64), (32, 4, 256, 256), (4, 4, 512, 512)])
def test_sharded_plans_partition_the_unsharded_plan(world, geom):
    n_ctx, n_gen, H, W = geom
    _, specs, n_lat, n_ctx_lat = _specs(n_ctx, n_gen, H, W)
    full = eng.build_plan(specs, n_lat, n_ctx_lat, H // 8, W // 8, "cpu")
    parts = [eng.build_plan(specs, n_lat, n_ctx_lat, H // 8, W // 8, "cpu", shard=(r, world)) for r in range(world)]
    for p in parts:      # global structures are replicated
        assert torch.equal(p.page_table, full.page_table) and torch.equal(p.k_code, full.k_code)
        assert torch.equal(p.k_tile_minmax, full.k_tile_minmax) and p.total_pages == full.total_pages
    for which in ("prefix", "step"):
        f = getattr(full, which)
        ph = [getattr(p, which) for p in parts]
        assert sum(x.rows for x in ph) == f.rows
        for s in range(len(specs)):
            # the rows of sequence s on all ranks, put back in sequence order (a row's K/V slot is unique and
            # ascending in the fresh plan), are exactly the unsharded plan's rows of that sequence
            fq0, fn = int(f.seqs[s, 0]), int(f.seqs[s, 1])
            pieces = {name: torch.cat([getattr(x, name)[int(x.seqs[s, 0]):int(x.seqs[s, 0]) + int(x.seqs[s, 1])] for x in ph])
                      for name in ("row_pos", "row_slot", "q_code", "kind", "arg_a", "arg_b")}
            order = torch.argsort(pieces["row_slot"])
            assert len(torch.unique(pieces["row_slot"])) == fn                # nobody's rows overlap
            for name, got in pieces.items():
                assert torch.equal(got[order], getattr(f, name)[fq0:fq0 + fn]), (which, name, s)
            for x in ph:       # inside a rank the rows of a sequence stay in ascending order (cheap chunk, then dear chunk)
                slots = x.row_slot[int(x.seqs[s, 0]):int(x.seqs[s, 0]) + int(x.seqs[s, 1])]
                assert bool((slots[1:] > slots[:-1]).all())
            sizes = [int(x.seqs[s, 1]) for x in ph]
            assert all(int(x.seqs[s, 2]) == int(f.seqs[s, 2]) for x in ph)   # every rank sees all keys
            if sum(sizes) >= 2 * eng.SHARD_ALIGN * world:
                # whole 128-row tiles everywhere; only the owner of the sequence's last rows carries the remainder
                assert sorted(n % eng.SHARD_ALIGN for n in sizes)[:-1] == [0] * (world - 1)
            else:
                assert max(sizes) - min(sizes) <= (1 if sum(sizes) < eng.SHARD_ALIGN * world else 2 * eng.SHARD_ALIGN - 1)


def _prefill_attention_tiles(ranges, block):
    """KV tiles a rank's attention CTAs (256 query rows each) walk for the given CONTEXT rows: context is
    frame-causal (`create_mask_frame_block_inference`: a context frame sees itself and the frames before it)."""
    rows = np.concatenate([np.arange(a, b) for a, b in ranges]) if ranges else np.zeros(0, int)
    total = 0
    for c0 in range(0, len(rows), 256):
        last = rows[min(c0 + 256, len(rows)) - 1]
        total += -(-((last // block + 1) * block) // 128)
    return total


@pytest.mark.parametrize("world,n,block", [(2, 4104, 1026), (4, 4104, 1026), (8, 4104, 1026), (2, 8256, 258), (8, 8256, 258)])
def test_shard_ranges_balance_the_causal_attention_cost_of_the_prefill(world, n, block):
    """Context rows of BASELINE configs[4] (4 frames of 1026 tokens) and configs[2] (32 frames of 258): with one
    contiguous chunk per rank the last rank's prefill attention walks far more KV tiles than the first (its queries
    are the last frames); cheapest-with-dearest chunk pairs even that out.  (The rows of the clip being denoised see
    every key whatever their frame -- the generated clip is bidirectional -- so for them only the placement of the
    remainder rows matters: next test.)"""
    contiguous = [_prefill_attention_tiles([eng.shard_rows(0, n, r, world)], block) for r in range(world)]
    paired = [_prefill_attention_tiles(eng.shard_ranges(0, n, r, world), block) for r in range(world)]
    assert sum(paired) <= sum(contiguous) + world and max(paired) < max(contiguous)
    assert max(paired) / (sum(paired) / world) < max(contiguous) / (sum(contiguous) / world) - 0.1


@pytest.mark.parametrize("world,n", [(2, 4104), (8, 4104), (2, 1032), (4, 1032)])
def test_remainder_rows_of_the_two_cfg_sequences_land_on_different_ranks(world, n):
    """A frame is 258 or 1026 tokens, so every sequence ends in 8 rows beyond the last 128-row tile: an attention CTA
    of their own per head (walking every KV tile) and a GEMM tail for whoever owns them.  `flip` (odd sequences)
    mirrors the deal: the conditional sequence's remainder goes to rank 0, the unconditional one's to the last rank
    (one contiguous chunk per rank put both on the last rank, and everybody waited for it at every layer)."""
    tail_owner = [r for r in range(world) if any(b == n for _, b in eng.shard_ranges(0, n, r, world))]
    tail_owner_flipped = [r for r in range(world) if any(b == n for _, b in eng.shard_ranges(0, n, r, world, flip=True))]
    assert tail_owner == [0] and tail_owner_flipped == [world - 1]
    for flip in (False, True):
        sizes = [sum(b - a for a, b in eng.shard_ranges(0, n, r, world, flip)) for r in range(world)]
        assert sorted(x % eng.SHARD_ALIGN for x in sizes) == [0] * (world - 1) + [n % eng.SHARD_ALIGN]


@pytest.mark.parametrize("geom", [(4, 4, 256, 256), (3, 2, 64, 96), (1, 1, 64, 64)])
def test_sequence_partition_gives_every_rank_whole_sequences(geom):
    """partition="sequences" with two ranks = one CFG branch per rank (CFG axis): rank 0 owns every
    row of the conditional sequence (context + clip), rank 1 every row of the unconditional one; the global
    structures (page table, key codes, latent numbering) stay those of the unsharded plan."""
    n_ctx, n_gen, H, W = geom
    _, specs, n_lat, n_ctx_lat = _specs(n_ctx, n_gen, H, W)
    full = eng.build_plan(specs, n_lat, n_ctx_lat, H // 8, W // 8, "cpu")
    parts = [eng.build_plan(specs, n_lat, n_ctx_lat, H // 8, W // 8, "cpu", shard=(r, 2), partition="sequences")
             for r in range(2)]
    for r, p in enumerate(parts):
        assert p.partition == "sequences" and p.shard == (r, 2)
        assert torch.equal(p.page_table, full.page_table) and torch.equal(p.k_code, full.k_code)
        assert torch.equal(p.lat_row0, full.lat_row0) and p.n_latents == full.n_latents
        for which in ("prefix", "step"):
            f, x = getattr(full, which), getattr(p, which)
            for s_ in range(2):
                q0, n = int(f.seqs[s_, 0]), int(f.seqs[s_, 1])
                assert int(x.seqs[s_, 1]) == (n if s_ == r else 0) and int(x.seqs[s_, 2]) == int(f.seqs[s_, 2])
            q0, n = int(f.seqs[r, 0]), int(f.seqs[r, 1])
            assert x.rows == n
            for name in ("row_pos", "row_slot", "q_code", "kind", "arg_a", "arg_b"):
                assert torch.equal(getattr(x, name), getattr(f, name)[q0:q0 + n]), (which, name)
    assert full.partition == "rows"
    with pytest.raises(ValueError, match="whole sequences"):
        eng.build_plan(specs, n_lat, n_ctx_lat, H // 8, W // 8, "cpu", shard=(0, 4), partition="sequences")
    with pytest.raises(ValueError, match="unknown partition"):
        eng.build_plan(specs, n_lat, n_ctx_lat, H // 8, W // 8, "cpu", shard=(0, 2), partition="heads")
    assert [eng.sequence_owner(s_, 8, 2) for s_ in range(8)] == [0] * 4 + [1] * 4      # batch: cond rows, then uncond rows


def test_shard_rows_covers_range():
    for lo, hi, w in ((0, 1032, 8), (1032, 2064, 3), (5, 7, 4), (0, 0, 2)):
        chunks = [eng.shard_rows(lo, hi, r, w) for r in range(w)]
        assert chunks[0][0] == lo and chunks[-1][1] == hi
        assert all(chunks[i][1] == chunks[i + 1][0] for i in range(w - 1))
    for lo, hi, w in ((0, 4104, 2), (0, 4104, 8), (1032, 2064, 2), (258, 8514, 8), (0, 1032, 8), (5, 7, 4), (0, 0, 2), (0, 520, 2)):
        for flip in (False, True):
            per_rank = [eng.shard_ranges(lo, hi, r, w, flip) for r in range(w)]
            cover = sorted(x for rs in per_rank for x in rs)
            assert all(b > a for a, b in cover) and all(rs == sorted(rs) and len(rs) <= 2 for rs in per_rank)
            if hi > lo:
                assert cover[0][0] == lo and cover[-1][1] == hi
                assert all(cover[i][1] == cover[i + 1][0] for i in range(len(cover) - 1))     # disjoint, complete
            assert per_rank == [eng.shard_ranges(lo, hi, w - 1 - r, w, not flip) for r in range(w)]


# ------------------------------------------------------------------------------------------------
# 1 GPU: virtual ranks in lockstep
# ------------------------------------------------------------------------------------------------
def _engine(w, dims, dev, peers=None):
    return eng.NextClipEngine(w, dims.hidden_size, dims.intermediate_size, dims.num_hidden_layers,
                              dims.num_attention_heads, dims.rms_norm_eps, dims.rope_theta, dev,
                              dims.pos_embed_max_size, 2, use_cuda_graph=False, peers=peers)


WIDE = synth.BackboneDims(num_hidden_layers=2)       # full width (32 heads x 96), two layers


@pytest.mark.gpu
@pytest.mark.parametrize("world,geom,dims,partition", [
    (2, (3, 2, 64, 96), synth.REDUCED, "rows"), (3, (2, 2, 64, 64), synth.REDUCED, "rows"),
    (4, (4, 4, 128, 128), synth.REDUCED, "rows"), (2, (4, 4, 256, 256), WIDE, "rows"),
    (3, (4, 4, 256, 256), WIDE, "rows"), (8, (2, 2, 128, 256), WIDE, "rows"),
    (2, (3, 2, 64, 96), synth.REDUCED, "sequences"), (2, (4, 4, 256, 256), WIDE, "sequences")])
def test_virtual_ranks_match_unsharded_engine_bit_exact(world, geom, dims, partition):
    from videogpt_b200 import ops, peer
    dev, bf = torch.device("cuda", 0), torch.bfloat16
    n_ctx, n_gen, H, W = geom
    sd = synth.init_state_dict(dims, seed=0, dtype=bf, with_pos_embed=False)
    w = eng.EngineWeights(sd, dims.num_hidden_layers, dev)
    _, specs, n_lat, n_ctx_lat = _specs(n_ctx, n_gen, H, W)
    lat = synth.synthetic_latents(n_ctx + n_gen, H, W, seed=42)
    ctx = torch.cat(lat[:n_ctx], 0).to(dev, bf)
    z0 = torch.cat(lat[n_ctx:] * 2, 0).to(dev, bf)

    ref = _engine(w, dims, dev)
    ref.set_plan(eng.build_plan(specs, n_lat, n_ctx_lat, H // 8, W // 8, dev))
    ref.prefill(ctx)
    members = peer.LocalPeerGroup.create(world, dev)
    ranks = [_engine(w, dims, dev, peers=m) for m in members]
    for r, e in enumerate(ranks):
        e.set_plan(eng.build_plan(specs, n_lat, n_ctx_lat, H // 8, W // 8, dev, shard=(r, world), partition=partition))
    eng.run_lockstep([e.prefill_steps(ctx) for e in ranks])
    torch.cuda.synchronize()

    def kv_as_expected(e, r):
        if partition == "rows":                       # every rank ends up with ALL K/V
            return torch.equal(e.kv, ref.kv)
        mine = ref.plan.page_table[r, :(specs[r].n_prefix + specs[r].n_active + 127) // 128].long()
        return torch.equal(e.kv[:, :, mine], ref.kv[:, :, mine])    # whole sequences: its own sequence's pages

    assert all(kv_as_expected(e, r) for r, e in enumerate(ranks))
    ref.z.copy_(z0)
    for e in ranks:
        e.z.copy_(z0)
    for step, t in enumerate((0.0, 0.3, 0.7)):
        for e in [ref] + ranks:
            e.t.fill_(t)
        ref.predict()
        eng.run_lockstep([e.predict_steps() for e in ranks])
        torch.cuda.synchronize()
        for r, e in enumerate(ranks):
            assert torch.equal(e.pred, ref.pred), f"step {step}: prediction differs"
            assert kv_as_expected(e, r)
        for e in [ref] + ranks:
            ops.cfg_euler(e.z, e.pred, True, True, 1.0 - t, 0.3, 1.5)
    assert all(torch.equal(e.z, ref.z) for e in ranks)


# ------------------------------------------------------------------------------------------------
# 2 GPUs: real peer memory
# ------------------------------------------------------------------------------------------------
def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, ret, geom=(3, 2, 64, 96), fresh_context=False, partition="rows"):
    import torch.distributed as dist
    from transformers import Phi3Config
    from videogpt_b200 import LVM, LVMScheduler, parallel_states as ps
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        dims = synth.REDUCED
        sd = synth.init_state_dict(dims, seed=0)

        def model():
            m = LVM(Phi3Config(**dims.phi3_kwargs()), device=dev)
            m.load_state_dict(sd)
            return m.to(torch.bfloat16).eval()

        (n_ctx, n_gen, H, W), steps = geom, 4
        d = po.frame_block_inputs(n_ctx, n_gen, H, W, True, 1)
        lat = [x.to(dev, torch.bfloat16) for x in synth.synthetic_latents(n_ctx + n_gen, H, W, seed=42)]
        mk = dict(input_ids=d["input_ids"].to(dev), input_img_latents=lat[:n_ctx],
                  input_image_sizes=d["input_image_sizes"], attention_mask=None, position_ids=d["position_ids"].to(dev),
                  denoise_image_sizes=d["denoise_image_sizes"], time_emb_inx=d["time_emb_inx"], img_cfg_scale=1.5,
                  use_img_cfg=True, use_kv_cache=False, offload_model=False, vae=None)
        single_model = model()
        single = {pt: LVMScheduler(steps)([x.clone() for x in lat[n_ctx:]] * 2, single_model.frame_block_forward_with_cfg,
                                          mk, prediction_type=pt) for pt in ("x1", "v")}
        ps.initialize_sequence_parallel_state(world, partition=partition)     # the reference's SP switch
        sp_model = model()
        out = {}
        for pt in ("x1", "v"):
            for rep in range(2):                               # second clip reuses plan + captured graph
                if fresh_context:      # new tensor objects, as for a new clip: the prefill runs again on the kept plan
                    mk = dict(mk, input_img_latents=[x.clone() for x in lat[:n_ctx]])
                got = LVMScheduler(steps)([x.clone() for x in lat[n_ctx:]] * 2, sp_model.frame_block_forward_with_cfg,
                                          mk, prediction_type=pt)
                torch.cuda.synchronize()
                out[f"{pt}{rep}"] = all(torch.equal(a, b) for a, b in zip(got, single[pt]))
        sp_model.engine().peers.check()
        plan = sp_model.engine().plan
        out["sharded"] = plan.shard == (rank, world) and plan.partition == partition
        if partition == "sequences":      # one CFG branch per rank: all rows of its own sequence, none of the other
            out["sharded"] &= plan.step.rows == specs_rows(d)[rank]
        ret[rank] = out
        dist.barrier()
        sp_model.engine().peers.close()
    finally:
        ps.destroy_sequence_parallel_group()
        dist.destroy_process_group()


def specs_rows(d):
    specs, _, _ = eng.frame_block_specs(d["input_ids"], d["position_ids"], d["input_image_sizes"],
                                        d["denoise_image_sizes"], d["time_emb_inx"])
    return [sp.n_active for sp in specs]


def _spawn_with_deadline(args, nprocs, seconds):
    """mp.spawn with a hard deadline: a stalled sequence-parallel group must cost seconds, not the NCCL watchdog's
    ten minutes (round 1)."""
    import time
    import torch.multiprocessing as mp
    ctx = mp.spawn(_worker, args=args, nprocs=nprocs, join=False)
    deadline = time.time() + seconds
    while not ctx.join(timeout=5):
        if time.time() > deadline:
            for p in ctx.processes:
                if p.is_alive():
                    p.kill()
            pytest.fail(f"sequence-parallel workers did not finish within {seconds} s")


@pytest.mark.gpu
@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
@pytest.mark.parametrize("geom,fresh", [((3, 2, 64, 96), False), ((3, 2, 64, 96), True), ((32, 4, 64, 64), True)])
def test_sequence_parallel_two_gpus_matches_single_gpu(geom, fresh):
    """Two real GPUs through LVM + LVMScheduler with CUDA graphs reproduce the single-GPU run bit for bit; the
    32-context-frame geometry is BASELINE configs[2]'s layout (8 context clips) at small frames; `fresh` = every
    clip brings new context tensors (prefill on a kept plan and graph, as bench.py does)."""
    import torch.multiprocessing as mp
    mgr = mp.Manager()
    ret = mgr.dict()
    _spawn_with_deadline((2, _free_port(), ret, geom, fresh), 2, 240)
    want = {"x10": True, "x11": True, "v0": True, "v1": True, "sharded": True}
    assert dict(ret) == {0: want, 1: want}


@pytest.mark.gpu
@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
@pytest.mark.parametrize("geom,fresh", [((3, 2, 64, 96), False), ((4, 4, 128, 128), True)])
def test_cfg_branch_pair_two_gpus_matches_single_gpu(geom, fresh):
    """One CFG branch per GPU (``initialize_sequence_parallel_state(2, partition="sequences")``): rank 0 runs the
    conditional sequence, rank 1 the unconditional one, each stores its half of the prediction into both ranks'
    buffers over NVLink and both apply the same update inside the step graph -- bit for bit the single-GPU latents,
    x1 and v, first clip and a second one on the kept plan and graph."""
    import torch.multiprocessing as mp
    mgr = mp.Manager()
    ret = mgr.dict()
    _spawn_with_deadline((2, _free_port(), ret, geom, fresh, "sequences"), 2, 240)
    want = {"x10": True, "x11": True, "v0": True, "v1": True, "sharded": True}
    assert dict(ret) == {0: want, 1: want}
