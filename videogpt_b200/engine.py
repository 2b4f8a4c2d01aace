"""Next-clip denoising engine: plans, paged KV cache, and the per-step kernel sequence.

This is the host side of the hot path.  It owns device memory through
PyTorch and enqueues the hand-written sm_100a kernels of ``libvgpt_b200.so``; it does no
arithmetic on tensors itself.

Design (B200-first, not the reference's control flow):

* A *sequence* is a step-invariant PREFIX (context frames) followed by ACTIVE rows (the
  clip being denoised).  Prefix rows never see active rows (mask closed form), so their K/V
  are computed once per clip (``prefill``) into a paged KV pool and every Euler step only
  runs the active rows (``predict``): the reference recomputes the context 50x and pads the
  unconditional row to full length (``LVM/scheduler.py:174``, ``LVM/processor.py:812-838``).
* The mask is never materialised: each token carries an integer code and the attention
  kernel evaluates ``code_q >= code_k`` (see ``processor.token_codes``).
* All rows of all sequences (cond / uncond / batch) are packed into one ``[M, hidden]``
  matrix, so every projection is one GEMM launch per layer.
* One step = 8 + 8 x layers kernel launches (the final RMSNorm, the final layer, unpatchify and -- in the sampler's
  fused loop -- the x1 -> v / CFG / Euler update are ONE kernel), captured once in a CUDA graph and replayed.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence

import numpy as np
import torch

from . import ops
from .ops import PAGE_TOKENS, ATTN_KV_TILE

INT_MAX = 2 ** 31 - 1
# Storage type of activations, weights and the KV pools.  The kernels only take bf16 (ops._req
# rejects anything else); the name exists so the CPU emulation in tests/emu_ops.py can run this
# module's host logic in fp32 against the fp32 oracle.
ACT_DTYPE = torch.bfloat16


# --------------------------------------------------------------------------------------------
# weights
# --------------------------------------------------------------------------------------------
class EngineWeights:
    """bf16 CUDA views of a state dict with the reference's names plus the
    packed copies the kernels want (gate/up block-interleaved for the SwiGLU epilogue)."""

    def __init__(self, sd: Dict[str, torch.Tensor], num_layers: int, device):
        def g(name):
            t = sd[name]
            if t.device != torch.device(device) or t.dtype != ACT_DTYPE or not t.is_contiguous():
                t = t.detach().to(device=device, dtype=ACT_DTYPE).contiguous()
            return t.detach()

        self.embed_tokens = g("llm.embed_tokens.weight")
        self.x_w, self.x_b = g("x_embedder.proj.weight"), g("x_embedder.proj.bias")
        self.cx_w, self.cx_b = g("input_x_embedder.proj.weight"), g("input_x_embedder.proj.bias")
        self.time_token = [g(f"time_token.mlp.{i}.{k}") for i in (0, 2) for k in ("weight", "bias")]
        self.t_embedder = [g(f"t_embedder.mlp.{i}.{k}") for i in (0, 2) for k in ("weight", "bias")]
        self.ada_w, self.ada_b = g("final_layer.adaLN_modulation.1.weight"), g("final_layer.adaLN_modulation.1.bias")
        self.final_w, self.final_b = g("final_layer.linear.weight"), g("final_layer.linear.bias")
        self.norm = g("llm.norm.weight")
        self.layers = []
        for n in range(num_layers):
            p = f"llm.layers.{n}."
            gate_up = g(p + "mlp.gate_up_proj.weight")
            self.layers.append(dict(
                ln1=g(p + "input_layernorm.weight"), qkv=g(p + "self_attn.qkv_proj.weight"),
                o=g(p + "self_attn.o_proj.weight"), ln2=g(p + "post_attention_layernorm.weight"),
                gate_up=ops.pack_gate_up(gate_up), down=g(p + "mlp.down_proj.weight")))
        self.pos_embed = sd.get("pos_embed")          # optional full table [1, max*max, h]


# --------------------------------------------------------------------------------------------
# plans
# --------------------------------------------------------------------------------------------
@dataclass
class SequenceSpec:
    """Host description of one sequence (all arrays int32, length n_prefix + n_active)."""
    n_prefix: int
    n_active: int
    positions: np.ndarray
    codes: np.ndarray
    kinds: np.ndarray
    arg_a: np.ndarray
    arg_b: np.ndarray
    latent_rows: List[tuple] = field(default_factory=list)   # (latent index, seq-relative first image row)
    # streaming rollout (rollout.py): the first n_cached prefix rows already have their K/V in the
    # pool (computed in an earlier round) and are not recomputed; `pages` names the physical pool
    # page of every logical page of this sequence (None: pages are dealt out by build_plan)
    n_cached: int = 0
    pages: Optional[np.ndarray] = None


@dataclass
class PhaseArrays:
    rows: int
    row_pos: torch.Tensor
    row_slot: torch.Tensor
    q_code: torch.Tensor
    kind: torch.Tensor
    arg_a: torch.Tensor
    arg_b: torch.Tensor
    seqs: torch.Tensor
    max_q_rows: int


@dataclass
class ClipPlan:
    specs: List[SequenceSpec]
    n_latents: int
    n_ctx_latents: int
    lat_h: int
    lat_w: int
    page_table: torch.Tensor
    k_code: torch.Tensor
    k_tile_minmax: torch.Tensor
    total_pages: int
    prefix: PhaseArrays
    step: PhaseArrays
    lat_row0: torch.Tensor
    max_pos: int
    shard: Optional[tuple] = None      # (rank, world) of a sharded plan (one rank of a peer group)
    partition: str = "rows"            # what the ranks of a sharded plan own: "rows" | "sequences"


def frame_block_specs(input_ids, position_ids, input_image_sizes, denoise_image_sizes, time_emb_inx):
    """Sequence specs of the frame-block (next-clip) layout from the reference's collated index
    dicts (``LVM/processor.py:964-1000``); row b of the batch becomes sequence b.

    Latent numbering follows the order the reference consumes them in
    (``LVM/model.py:436-453``): context latents and noisy latents are global running counters
    over rows."""
    ids = input_ids.cpu().numpy()
    pos = position_ids.cpu().numpy()
    B, L = ids.shape
    specs, ctx_counter, lat_counter = [], 0, 0
    for b in range(B):
        ctx_ranges = list(input_image_sizes.get(b, []))
        gen_ranges = list(denoise_image_sizes.get(b, []))
        if not gen_ranges:
            raise ValueError(f"row {b}: no frames to denoise")
        t_inx = list(time_emb_inx.get(b, []))
        if len(t_inx) != len(gen_ranges):
            raise ValueError(f"row {b}: time_emb_inx does not match denoise_image_sizes")
        n_frames = len(ctx_ranges) + len(gen_ranges)
        first = (ctx_ranges or gen_ranges)[0][0]
        pad = first - 1 if ctx_ranges else first - 2          # processor.py:508-511
        token_l = gen_ranges[-1][1] - pad
        if gen_ranges[-1][1] != L or token_l % n_frames != 0:
            raise ValueError(f"row {b}: frame blocks must have equal length and end at the sequence end")
        bl = token_l // n_frames
        n_ctx, n_gen = len(ctx_ranges), len(gen_ranges)
        T = token_l
        kinds = np.full(T, ops.ROW_TOKEN, np.int32)
        arg_a = ids[b, pad:].astype(np.int32).copy()
        arg_b = np.zeros(T, np.int32)
        for s, e in ctx_ranges:
            kinds[s - pad:e - pad] = ops.ROW_CONTEXT_PATCH
            arg_a[s - pad:e - pad] = ctx_counter
            arg_b[s - pad:e - pad] = np.arange(e - s)
            ctx_counter += 1
        latent_rows = []
        for (s, e), ti in zip(gen_ranges, t_inx):
            if ti != s - 1:
                raise ValueError(f"row {b}: time slot must precede its image tokens")
            kinds[ti - pad] = ops.ROW_TIME
            arg_a[ti - pad] = lat_counter
            kinds[s - pad:e - pad] = ops.ROW_NOISY_PATCH
            arg_a[s - pad:e - pad] = lat_counter
            arg_b[s - pad:e - pad] = np.arange(e - s)
            latent_rows.append((lat_counter, s - pad))
            lat_counter += 1
        # codes: 4 * min(frame, n_ctx) + rank  (module docstring of processor.py)
        i = np.arange(T)
        f, o = i // bl, i % bl
        gen = f >= n_ctx
        r = np.where(gen, np.minimum(o, 2), np.where(o == 0, 0, np.where(o == bl - 1, 2, 1)))
        codes = (4 * np.minimum(f, n_ctx) + r).astype(np.int32)
        specs.append(SequenceSpec(n_prefix=n_ctx * bl, n_active=n_gen * bl,
                                  positions=pos[b, pad:].astype(np.int32), codes=codes, kinds=kinds,
                                  arg_a=arg_a, arg_b=arg_b, latent_rows=latent_rows))
    return specs, lat_counter, ctx_counter


def single_frame_specs(input_ids, position_ids, input_image_sizes, n_out_tokens: int):
    """Sequence specs of the one-frame-at-a-time layout of ``pipeline.__call__``
    (``LVMCollator.__call__``, ``LVM/processor.py:943-962``; assembled in ``LVM.forward``,
    ``LVM/model.py:345-362``): row b = [pad | condition tokens with context-image ranges | time
    token | n_out_tokens noisy image tokens].  Prefix = condition tokens, active = time + image.

    Codes reproduce ``create_mask`` + ``adjust_attention_for_input_images``
    (processor.py:536-573, 776-781): condition tokens are causal (code = index), the tokens of one
    context image share the code of its last token (bidirectional block), the time token is
    causal, output-image tokens carry the largest code (they see everything; nothing earlier
    sees them)."""
    ids = input_ids.cpu().numpy()
    pos = position_ids.cpu().numpy()
    B, Lt = ids.shape
    L = Lt + 1 + n_out_tokens
    assert pos.shape == (B, L), "position_ids must cover condition + time + image tokens"
    pads = (Lt + n_out_tokens + 1) - (pos.max(axis=1) + 1)          # create_position: pos = idx - pad
    specs, ctx_counter = [], 0
    for b in range(B):
        pad = int(pads[b])
        t_real = Lt - pad
        T = t_real + 1 + n_out_tokens
        kinds = np.full(T, ops.ROW_TOKEN, np.int32)
        arg_a = np.zeros(T, np.int32)
        arg_b = np.zeros(T, np.int32)
        arg_a[:t_real] = ids[b, pad:]
        codes = np.arange(T, dtype=np.int32)
        for s_, e_ in input_image_sizes.get(b, []):
            kinds[s_ - pad:e_ - pad] = ops.ROW_CONTEXT_PATCH
            arg_a[s_ - pad:e_ - pad] = ctx_counter
            arg_b[s_ - pad:e_ - pad] = np.arange(e_ - s_)
            codes[s_ - pad:e_ - pad] = e_ - pad - 1
            ctx_counter += 1
        kinds[t_real] = ops.ROW_TIME
        arg_a[t_real] = b
        kinds[t_real + 1:] = ops.ROW_NOISY_PATCH
        arg_a[t_real + 1:] = b
        arg_b[t_real + 1:] = np.arange(n_out_tokens)
        codes[t_real + 1:] = T
        specs.append(SequenceSpec(n_prefix=t_real, n_active=1 + n_out_tokens,
                                  positions=pos[b, pad:].astype(np.int32), codes=codes, kinds=kinds,
                                  arg_a=arg_a, arg_b=arg_b, latent_rows=[(b, t_real + 1)]))
    return specs, B, ctx_counter


def codes_dense_mask(spec_codes: np.ndarray, pad: int) -> np.ndarray:
    """Dense bool mask (with ``pad`` left-pad tokens) implied by a code array -- host mirror of
    ``vgpt_mask_from_codes``, used to validate a caller-supplied ``attention_mask``."""
    T = len(spec_codes)
    L = T + pad
    qc = np.concatenate([np.full(pad, INT_MAX, np.int64), spec_codes.astype(np.int64)])
    kc = np.concatenate([np.full(pad, INT_MAX - 1, np.int64), spec_codes.astype(np.int64)])
    return qc[:, None] >= kc[None, :]


SHARD_ALIGN = 128      # attention query tile; GEMM tiles and attention CTAs are 256 rows (pairs of 128)


def shard_rows(lo: int, hi: int, rank: int, world: int):
    """Rows [lo, hi) dealt to ``world`` ranks as ONE contiguous chunk each, like the reference's
    ``input_emb[:, r*L/P:(r+1)*L/P]`` (``LVM/model.py:459-464``) but without its divisibility
    requirement.  When every rank gets at least one 128-row tile the chunk boundaries are rounded down
    to multiples of 128 rows (the last rank takes the remainder): a shard then starts on a tile boundary
    of the attention kernel and the remainder rows meet the GEMM as one tail.  Otherwise sizes differ
    by at most one.  (``shard_ranges`` falls back to this for short ranges.)"""
    n = hi - lo
    if n >= SHARD_ALIGN * world:
        cut = lambda r: n if r >= world else (n * r // world) // SHARD_ALIGN * SHARD_ALIGN
    else:
        cut = lambda r: (n * r) // world
    return lo + cut(rank), lo + cut(rank + 1)


def shard_ranges(lo: int, hi: int, rank: int, world: int, flip: bool = False):
    """Rows [lo, hi) of one sequence dealt to ``world`` ranks, balanced for the attention cost: every row-wise op
    is independent of the partition and attention sees all keys through the token codes, so a rank may own ANY
    subset of the rows and gets the same numbers (tests/test_sequence_parallel.py, bit for bit).

    Context rows are frame-causal (a context frame sees itself and the frames before it), so with one contiguous
    chunk per rank the last rank's PREFILL attention walks up to twice the KV tiles of the average rank and everybody
    waits for it at every layer's barrier.  Here the range is cut into ``2 * world`` chunks on 256-row boundaries
    (one attention CTA = one GEMM tile row; 128 when the range is too short for that) and rank r takes chunk r and
    chunk ``2 * world - 1 - r`` -- the cheapest with the dearest.  (The rows of the clip being denoised see every key
    whatever their frame; for them the deal changes nothing but the next point.)  The remainder rows (a frame is 258
    or 1026 tokens: 8 rows at the end of every sequence) cost their owner an attention CTA of its own per head that
    walks every KV tile, and a GEMM tail; ``flip`` (odd sequences) mirrors the deal, so the conditional sequence's
    remainder lands on rank 0 and the unconditional one's on the last rank instead of both on the same.  Returns
    ascending (start, end) ranges."""
    n = hi - lo
    unit = next((u for u in (2 * SHARD_ALIGN, SHARD_ALIGN) if n >= u * 2 * world), 0)
    r = world - 1 - rank if flip else rank
    if not unit:
        a, b = shard_rows(lo, hi, r, world)
        return [(a, b)] if b > a else []
    cut = lambda i: n if i >= 2 * world else (n * i // (2 * world)) // unit * unit
    first, second = (cut(r), cut(r + 1)), (cut(2 * world - 1 - r), cut(2 * world - r))
    if first[1] == second[0]:                  # the two middle chunks are neighbours
        return [(lo + first[0], lo + second[1])]
    return [(lo + a, lo + b) for a, b in (first, second) if b > a]


def sequence_owner(s: int, n_seqs: int, world: int) -> int:
    """Rank that owns sequence ``s`` of ``n_seqs`` under ``partition="sequences"``: contiguous blocks, so with the
    reference's batch order ``[conditional rows | unconditional rows]`` (``LVM/pipeline.py:441-448``) and two ranks,
    rank 0 runs every conditional sequence (context + clip) and rank 1 every unconditional one -- the CFG-branch
    axis of SURVEY.md 8(e)."""
    if n_seqs % world:
        raise ValueError(f"{n_seqs} sequences cannot be dealt to {world} ranks as whole sequences")
    return s * world // n_seqs


def build_plan(specs: Sequence[SequenceSpec], n_latents: int, n_ctx_latents: int, lat_h: int,
               lat_w: int, device, shard: Optional[tuple] = None, max_pages: Optional[int] = None,
               pool_pages: Optional[int] = None, partition: str = "rows") -> ClipPlan:
    """``shard=(rank, world)``: plan of one rank of a peer group.  ``partition="rows"`` (sequence
    parallelism): the page table, key codes and tile classification describe the WHOLE sequences
    (every rank holds all K/V), the per-phase row arrays only this rank's chunk of every sequence.
    ``partition="sequences"`` (CFG branches on rank pairs): a rank owns whole sequences
    (``sequence_owner``) and no rows of the others; K/V never travel, only the prediction does.

    ``max_pages`` / ``pool_pages``: fixed capacities of the page table (logical pages per
    sequence) and of the pool, for plans that are refreshed in place round after round
    (``NextClipEngine.set_plan(keep_kv=True)``); specs may then carry their own physical pages and
    a cached prefix."""
    S = len(specs)
    if shard is not None:
        s_rank, s_world = shard
        if not (0 <= s_rank < s_world):
            raise ValueError(f"bad shard {shard}")
        if partition not in ("rows", "sequences"):
            raise ValueError(f"unknown partition {partition!r}")
        if partition == "sequences":
            sequence_owner(0, S, s_world)          # raises when whole sequences cannot be dealt evenly
    for sp in specs:
        T = sp.n_prefix + sp.n_active
        assert all(len(a) == T for a in (sp.positions, sp.codes, sp.kinds, sp.arg_a, sp.arg_b))
        if sp.n_prefix and sp.n_active:
            # caching the prefix is exact only if no prefix query can see an active key (rows whose
            # K/V are already cached are no queries here; their codes only describe them as keys)
            if sp.n_cached < sp.n_prefix and \
                    int(sp.codes[sp.n_cached:sp.n_prefix].max()) >= int(sp.codes[sp.n_prefix:].min()):
                raise ValueError("prefix rows can see active rows: this mask cannot be prefix-cached")
    pages = [(sp.n_prefix + sp.n_active + PAGE_TOKENS - 1) // PAGE_TOKENS for sp in specs]
    if max_pages is None:
        max_pages = max(pages)
    elif max(pages) > max_pages:
        raise ValueError(f"a sequence needs {max(pages)} pages, the page table holds {max_pages}")
    page_table = np.zeros((S, max_pages), np.int32)
    explicit = [sp.pages is not None for sp in specs]
    if any(explicit) != all(explicit):
        raise ValueError("either every sequence names its physical pages or none does")
    base = 0
    for s, n in enumerate(pages):
        if explicit[s]:
            if len(specs[s].pages) < n:
                raise ValueError(f"sequence {s}: {n} logical pages, {len(specs[s].pages)} physical pages given")
            page_table[s, :n] = specs[s].pages[:n]
            base = max(base, int(np.max(specs[s].pages[:n])) + 1)
        else:
            page_table[s, :n] = np.arange(base, base + n)
            base += n
    if all(explicit):
        used = np.concatenate([page_table[s, :n] for s, n in enumerate(pages)])
        if len(np.unique(used)) != len(used):
            raise ValueError("two logical pages share a physical page")
    if pool_pages is not None:
        if base > pool_pages:
            raise ValueError(f"plan needs physical page {base - 1}, the pool holds {pool_pages}")
        base = pool_pages
    for sp in specs:
        if not (0 <= sp.n_cached <= sp.n_prefix):
            raise ValueError("n_cached must lie inside the prefix")
    tiles_per_page = PAGE_TOKENS // ATTN_KV_TILE
    max_k_tiles = max_pages * tiles_per_page
    k_code = np.full((S, max_pages * PAGE_TOKENS), INT_MAX, np.int32)
    minmax = np.zeros((S, max_k_tiles, 2), np.int32)
    minmax[:, :, 0] = INT_MAX
    minmax[:, :, 1] = INT_MAX
    for s, sp in enumerate(specs):
        T = sp.n_prefix + sp.n_active
        k_code[s, :T] = sp.codes
        for t in range((T + ATTN_KV_TILE - 1) // ATTN_KV_TILE):
            seg = sp.codes[t * ATTN_KV_TILE:min(T, (t + 1) * ATTN_KV_TILE)]
            minmax[s, t] = (seg.min(), seg.max())

    def phase(which: str) -> PhaseArrays:
        pos, slot, qc, kd, aa, ab, seqs = [], [], [], [], [], [], []
        row0, max_q = 0, 0
        for s, sp in enumerate(specs):
            lo, hi = (sp.n_cached, sp.n_prefix) if which == "prefix" else (sp.n_prefix, sp.n_prefix + sp.n_active)
            kv_len = hi                        # every key up to the end of this phase's rows
            ranges = [(lo, hi)]
            if shard is not None and partition == "rows":
                ranges = shard_ranges(lo, hi, s_rank, s_world, flip=bool(s & 1))
            elif shard is not None and sequence_owner(s, S, s_world) != s_rank:
                ranges = []                        # somebody else's sequence: present (page table, codes), no rows here
            logical = np.concatenate([np.arange(a, b) for a, b in ranges] + [np.zeros(0, np.int64)]).astype(np.int64)
            n = len(logical)
            pos.append(sp.positions[logical])
            slot.append(page_table[s, logical // PAGE_TOKENS] * PAGE_TOKENS + logical % PAGE_TOKENS)
            qc.append(sp.codes[logical]); kd.append(sp.kinds[logical]); aa.append(sp.arg_a[logical]); ab.append(sp.arg_b[logical])
            seqs.append([row0, n, kv_len, 0])
            row0 += n
            max_q = max(max_q, n)
        cat = lambda xs: torch.from_numpy(np.concatenate(xs).astype(np.int32) if xs else np.zeros(0, np.int32)).to(device)
        return PhaseArrays(rows=row0, row_pos=cat(pos), row_slot=cat(slot), q_code=cat(qc), kind=cat(kd),
                           arg_a=cat(aa), arg_b=cat(ab),
                           seqs=torch.tensor(seqs, dtype=torch.int32, device=device), max_q_rows=max_q)

    prefix, step = phase("prefix"), phase("step")
    lat_row0 = np.zeros(max(n_latents, 1), np.int32)
    row0 = 0
    for sp in specs:
        for lat, first in sp.latent_rows:
            lat_row0[lat] = row0 + first - sp.n_prefix
        row0 += sp.n_active
    max_pos = int(max(int(sp.positions.max()) for sp in specs)) + 1
    return ClipPlan(specs=list(specs), n_latents=n_latents, n_ctx_latents=n_ctx_latents, lat_h=lat_h,
                    lat_w=lat_w, page_table=torch.from_numpy(page_table).to(device),
                    k_code=torch.from_numpy(k_code).to(device),
                    k_tile_minmax=torch.from_numpy(minmax).to(device), total_pages=base,
                    prefix=prefix, step=step, lat_row0=torch.from_numpy(lat_row0).to(device),
                    max_pos=max_pos, shard=shard, partition=partition if shard is not None else "rows")


# --------------------------------------------------------------------------------------------
# engine
# --------------------------------------------------------------------------------------------
class NextClipEngine:
    def __init__(self, weights: EngineWeights, hidden_size: int, intermediate_size: int, num_layers: int,
                 num_heads: int, rms_eps: float, rope_theta: float, device, pos_embed_max_size: int = 192,
                 patch_size: int = 2, use_cuda_graph: bool = True, peers=None):
        """``peers``: a ``peer.PeerGroup`` (one process per GPU) or ``peer.LocalPeerGroup`` member
        when this engine is one rank of a sequence-parallel group; plans must then be built with
        ``shard=(peers.rank, peers.world)``."""
        if not torch.cuda.is_available():
            raise RuntimeError("videogpt_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
        self.w = weights
        self.hs, self.inter, self.L, self.H = hidden_size, intermediate_size, num_layers, num_heads
        self.D = hidden_size // num_heads
        self.eps, self.theta = rms_eps, rope_theta
        self.device = torch.device(device)
        self.pos_max, self.patch = pos_embed_max_size, patch_size
        self.use_cuda_graph = use_cuda_graph
        self.peers = peers
        self._kv_shared = self._pred_shared = None
        # True when the caller guarantees that every latent carries the same timestep (the fused
        # sampler loop: scheduler.py:171 builds `timesteps` from one sigma): the three small MLPs
        # then run on ONE row and the result is replicated -- identical numbers, 1/n of the work.
        self.uniform_t = False
        # Set by the sampler's fused loop (scheduler.run_prepared): (use_cfg, x1_mode) -> the step also applies its
        # x1 -> v / CFG / Euler update to self.z, reading the step scalars from self.scalars -- inside the final-layer
        # kernel on one GPU, as vgpt_cfg_euler behind the last barrier in a peer group; None: predict() only produces
        # self.pred (the S2 callback seam)
        self.euler_mode = None
        self._graph_key = None
        self.plan: Optional[ClipPlan] = None
        # diagnostic tap (tools/parity_floor.py): a list that receives a copy of the hidden rows after every
        # decoder layer of eager (non-graph) passes; None in production
        self.layer_tap: Optional[list] = None
        self._graph = None
        self._rope_tab = None
        self.rope_reserve = 0            # positions to provision the RoPE table for (rollouts grow)
        half = 128
        self._t_freqs = torch.exp(-math.log(10000.0) * torch.arange(half, dtype=torch.float32) / half).to(self.device)
        self._inv_freq = (1.0 / (rope_theta ** (torch.arange(0, self.D, 2, dtype=torch.int64).float() / self.D))).to(self.device)

    # ---- setup ---------------------------------------------------------------------------------
    def _refresh_in_place(self, new: ClipPlan) -> bool:
        """Same shapes as the current plan (a later round of a rollout, the next clip of a stream):
        overwrite the plan's device arrays instead of replacing them, so workspaces, K/V pool and the
        captured CUDA graph -- which holds these arrays' addresses -- all stay valid."""
        old = self.plan
        if old is None or self.peers is not None or old.shard is not None or new.shard is not None:
            return False
        scal = lambda p: (p.n_latents, p.n_ctx_latents, p.lat_h, p.lat_w, p.total_pages, p.prefix.rows, p.step.rows,
                          p.prefix.max_q_rows, p.step.max_q_rows)
        if scal(old) != scal(new) or self._rope_tab is None or new.max_pos > self._rope_tab.shape[0]:
            return False
        pairs = [(getattr(old, k), getattr(new, k)) for k in ("page_table", "k_code", "k_tile_minmax", "lat_row0")]
        for ph in ("prefix", "step"):
            o, n = getattr(old, ph), getattr(new, ph)
            pairs += [(getattr(o, k), getattr(n, k)) for k in ("row_pos", "row_slot", "q_code", "kind", "arg_a", "arg_b", "seqs")]
        if any(a.shape != b.shape for a, b in pairs):
            return False
        for a, b in pairs:
            a.copy_(b)
        old.specs, old.max_pos = new.specs, new.max_pos
        self.prefilled = False
        return True

    def set_plan(self, plan: ClipPlan, keep_kv: bool = False):
        """``keep_kv``: the K/V pool carries rows cached by an earlier plan (``SequenceSpec.n_cached``):
        keep its contents -- and, when nothing but the contents of the plan's arrays changed, keep
        everything (``_refresh_in_place``)."""
        if keep_kv and self._refresh_in_place(plan):
            return
        old_kv = getattr(self, "kv", None) if keep_kv and self.peers is None else None
        self.plan = plan
        self._graph = None
        dev, bf = self.device, ACT_DTYPE
        rows = max(plan.prefix.rows, plan.step.rows, 1)
        self.hidden = torch.empty(rows, self.hs, device=dev, dtype=bf)
        self.xn = torch.empty(rows, self.hs, device=dev, dtype=bf)
        self.qkv = torch.empty(rows, 3 * self.hs, device=dev, dtype=bf)
        self.attn = torch.empty(rows, self.hs, device=dev, dtype=bf)
        self.mlp_h = torch.empty(rows, self.inter, device=dev, dtype=bf)
        # paged KV pools: [layer][k|v][page][H][128][D]
        n = max(plan.n_latents, 1)
        self.z = torch.zeros(n, 4, plan.lat_h, plan.lat_w, device=dev, dtype=bf)
        kv_shape = (self.L, 2, plan.total_pages, self.H, PAGE_TOKENS, self.D)
        if self.peers is None:
            if plan.shard is not None:
                raise ValueError("a row-sharded plan needs an engine with a peer group")
            if old_kv is not None and tuple(old_kv.shape) == kv_shape:
                self.kv = old_kv              # physical pages keep their contents
            else:
                if any(sp.n_cached for sp in plan.specs):
                    raise ValueError("the plan counts on cached K/V rows but the pool is being (re)allocated")
                self.kv = torch.zeros(kv_shape, device=dev, dtype=bf)
            self.pred = torch.zeros_like(self.z)
            self._kv_ptrs = None
        else:
            # sequence parallel: pools and prediction live in peer-mapped memory; every rank holds
            # ALL K/V (the producers store into every peer) and the full prediction
            if plan.shard != (self.peers.rank, self.peers.world):
                raise ValueError(f"plan shard {plan.shard} does not match the peer group "
                                 f"({self.peers.rank}, {self.peers.world})")
            esize = self.z.element_size()
            kv_bytes, pred_bytes = esize * math.prod(kv_shape), esize * self.z.numel()
            if self._kv_shared is None or self._kv_shared.local.numel() < kv_bytes:
                if any(sp.n_cached for sp in plan.specs):
                    raise ValueError("the plan counts on cached K/V rows but the pool is being (re)allocated")
                if self._kv_shared is not None:          # outgrown: release it on every rank before growing
                    self.kv = None
                    self.peers.free(self._kv_shared)
                self._kv_shared = self.peers.alloc(kv_bytes)
            if self._pred_shared is None or self._pred_shared.local.numel() < pred_bytes:
                if self._pred_shared is not None:
                    self.pred = None
                    self.peers.free(self._pred_shared)
                self._pred_shared = self.peers.alloc(pred_bytes)
            self.kv = self._kv_shared.local[:kv_bytes].view(bf).view(kv_shape)
            self.pred = self._pred_shared.local[:pred_bytes].view(bf).view(self.z.shape)
            pool_bytes = esize * math.prod(kv_shape[2:])
            self._kv_ptrs = [(self._kv_shared.ptr_array((2 * li) * pool_bytes),
                              self._kv_shared.ptr_array((2 * li + 1) * pool_bytes)) for li in range(self.L)]
            self._pred_ptrs = self._pred_shared.ptr_array()
        self.ctx = torch.zeros(max(plan.n_ctx_latents, 1), 4, plan.lat_h, plan.lat_w, device=dev, dtype=bf)
        # per-step inputs in ONE buffer so that the sampler sets them with one copy per step:
        # [timestep of every latent | 1 - sigma, d sigma, guidance]
        self.step_inputs = torch.zeros(n + 3, device=dev, dtype=torch.float32)
        self.t = self.step_inputs[:n]
        self.scalars = self.step_inputs[n:n + 3]
        self.vel = torch.zeros_like(self.z)
        self.t_sin = torch.empty(n, 256, device=dev, dtype=bf)
        self.t_h1 = torch.empty(n, self.hs, device=dev, dtype=bf)
        self.time_tokens = torch.empty(n, self.hs, device=dev, dtype=bf)
        self.t_emb = torch.empty(n, self.hs, device=dev, dtype=bf)
        self.mod = torch.empty(n, 2 * self.hs, device=dev, dtype=bf)
        self.pos_rows = self._pos_rows(plan.lat_h, plan.lat_w)
        if self._rope_tab is None or self._rope_tab.shape[0] < plan.max_pos:
            self._rope_tab = ops.rope_table(self._inv_freq, max(plan.max_pos, self.rope_reserve, 1), self.D)
        self.prefilled = False
        if self.peers is not None:
            self.peers.host_barrier()         # every rank has finished allocating

    def _pos_rows(self, lat_h: int, lat_w: int) -> torch.Tensor:
        """``cropped_pos_embed`` (LVM/model.py:268-289) as bf16 rows ``[tokens, hidden]``: gathered
        from the model's persistent buffer when it is present, else computed for the crop."""
        hh, ww = lat_h // self.patch, lat_w // self.patch
        if hh > self.pos_max:
            raise ValueError(f"Height ({hh}) cannot be greater than `pos_embed_max_size`: {self.pos_max}.")
        if ww > self.pos_max:
            raise ValueError(f"Width ({ww}) cannot be greater than `pos_embed_max_size`: {self.pos_max}.")
        if self.w.pos_embed is not None:
            top, left = (self.pos_max - hh) // 2, (self.pos_max - ww) // 2
            pe = self.w.pos_embed.reshape(self.pos_max, self.pos_max, -1)[top:top + hh, left:left + ww]
            return pe.reshape(hh * ww, -1).to(device=self.device, dtype=ACT_DTYPE).contiguous()
        from .synth import cropped_pos_embed_rows
        return cropped_pos_embed_rows(self.hs, lat_h, lat_w, self.patch, self.pos_max).to(
            device=self.device, dtype=ACT_DTYPE).contiguous()

    # ---- kernel sequences ----------------------------------------------------------------------
    def _time_embeddings(self, n: int):
        """time_token / t_embedder MLPs and the adaLN modulation (LVM/model.py:420, 480, 80)."""
        w = self.w
        full_n = n
        if self.uniform_t:
            n = 1
        ops.timestep_sinusoid(self.t[:n], self._t_freqs, self.t_sin[:n])
        ops.linear_small(self.t_sin[:n], w.time_token[0], w.time_token[1], post_silu=True, out=self.t_h1[:n])
        ops.linear_small(self.t_h1[:n], w.time_token[2], w.time_token[3], out=self.time_tokens[:n])
        ops.linear_small(self.t_sin[:n], w.t_embedder[0], w.t_embedder[1], post_silu=True, out=self.t_h1[:n])
        ops.linear_small(self.t_h1[:n], w.t_embedder[2], w.t_embedder[3], out=self.t_emb[:n])
        ops.linear_small(self.t_emb[:n], w.ada_w, w.ada_b, pre_silu=True, out=self.mod[:n])
        if full_n > n:                       # replicate row 0 (device-to-device copies, graph capturable)
            self.time_tokens[1:full_n].copy_(self.time_tokens[0:1].expand(full_n - 1, -1))
            self.mod[1:full_n].copy_(self.mod[0:1].expand(full_n - 1, -1))

    def _assemble(self, ph: PhaseArrays):
        w = self.w
        ops.embed_assemble(self.hidden[:ph.rows], ph.kind, ph.arg_a, ph.arg_b, w.embed_tokens, self.time_tokens,
                           self.z, self.ctx, self.plan.lat_h, self.plan.lat_w, w.x_w, w.x_b, w.cx_w, w.cx_b,
                           self.pos_rows)

    def _layers(self, ph: PhaseArrays, kv_only_last: bool):
        """Phi3DecoderLayer x L (transformers 4.47.1) on the packed rows of one phase.  Generator:
        yields at every point where, in a sequence-parallel group, the peers' stores must have
        landed before the next kernel (once per layer, after the K/V append)."""
        n, plan = ph.rows, self.plan
        hidden, xn, qkv, attn, mlp_h = (self.hidden[:n], self.xn[:n], self.qkv[:n], self.attn[:n], self.mlp_h[:n])
        scale = 1.0 / math.sqrt(self.D)
        # K/V travel only when ranks share sequences; a rank that owns whole sequences keeps its K/V to itself
        share_kv = self.peers is not None and plan.partition == "rows"
        for li, lw in enumerate(self.w.layers):
            if n:
                ops.rmsnorm(hidden, lw["ln1"], self.eps, out=xn)
                ops.gemm(xn, lw["qkv"], out=qkv)
                if not share_kv:
                    ops.rope_kv_append(qkv, ph.row_pos, ph.row_slot, self._rope_tab, self.kv[li, 0], self.kv[li, 1],
                                       self.H, self.D)
                else:                 # RoPE + K/V append + all-gather: stores into every rank's pool
                    ops.rope_kv_append_peers(qkv, ph.row_pos, ph.row_slot, self._rope_tab, self._kv_ptrs[li][0],
                                             self._kv_ptrs[li][1], self.peers.world, self.H, self.D)
            if share_kv:
                yield "kv"
            if kv_only_last and li == self.L - 1:
                break                     # prefix rows: nothing after the last K/V append is read
            if not n:
                continue
            ops.attention(qkv[:, :self.hs], attn, self.kv[li, 0], self.kv[li, 1], plan.page_table, ph.seqs,
                          ph.max_q_rows, ph.q_code, plan.k_code, plan.k_tile_minmax, self.H, self.D, scale)
            ops.gemm(attn, lw["o"], out=hidden, residual=hidden, epilogue=ops.EPI_RESIDUAL)
            ops.rmsnorm(hidden, lw["ln2"], self.eps, out=xn)
            ops.gemm(xn, lw["gate_up"], out=mlp_h, epilogue=ops.EPI_SWIGLU)
            ops.gemm(mlp_h, lw["down"], out=hidden, residual=hidden, epilogue=ops.EPI_RESIDUAL)
            if self.layer_tap is not None:
                self.layer_tap.append(hidden.clone())

    def _drive(self, gen):
        """Run a kernel-sequence generator; at its sync points enqueue the cross-GPU barrier."""
        for _ in gen:
            self.peers.barrier()

    def prefill_steps(self, ctx_latents: Optional[torch.Tensor] = None):
        """Generator form of ``prefill`` (``run_lockstep`` interleaves the virtual ranks of a
        ``LocalPeerGroup`` with it)."""
        plan = self.plan
        if ctx_latents is not None and plan.n_ctx_latents:
            self.ctx.copy_(ctx_latents.to(self.ctx.dtype).reshape(self.ctx.shape))
        if self.peers is not None:
            # clip boundary: from here to the last kernel of the clip this rank's host never blocks
            # in the driver (no allocation, no capture) while a peer may be spinning on it
            self.peers.host_barrier()
        if any(sp.n_prefix for sp in plan.specs):
            if plan.prefix.rows:
                self._assemble(plan.prefix)
            yield from self._layers(plan.prefix, kv_only_last=True)
        self.prefilled = True

    def prefill(self, ctx_latents: Optional[torch.Tensor] = None):
        """Compute and cache K/V of every prefix (context) row.  ``ctx_latents``: [n_ctx, 4, h, w]."""
        self._drive(self.prefill_steps(ctx_latents))

    def _predict_kernels(self):
        plan = self.plan
        st = plan.step
        if self.peers is not None and plan.partition == "sequences":
            # The only other barrier of such a step is the one after the prediction stores, and a peer's stores of
            # step i+1 must not land in this rank's `pred` before its update of step i has read it.  (Row-sharded
            # steps have a barrier per layer in between.)
            yield "start"
        self._time_embeddings(plan.n_latents)
        if st.rows:
            self._assemble(st)
        yield from self._layers(st, kv_only_last=False)
        # llm.norm + FinalLayer + unpatchify (+ the scheduler update in the fused loop): one kernel on the raw residual stream
        n_half = plan.n_latents // 2 if self.euler_mode is not None and self.euler_mode[0] else plan.n_latents
        if self.peers is None:
            euler = None
            if self.euler_mode is not None:
                use_cfg, x1_mode = self.euler_mode
                euler = (self.z, self.scalars, use_cfg, x1_mode, self.vel[:n_half])
            ops.final_layer(self.hidden[:st.rows], plan.lat_row0, self.mod[:plan.n_latents], self.w.final_w,
                            self.w.final_b, self.pred, norm_weight=self.w.norm, rms_eps=self.eps, euler=euler)
        else:                             # prediction stored into every rank's pred buffer
            if st.rows:
                ops.final_layer_rows(self.hidden[:st.rows], st.kind, st.arg_a, st.arg_b, self.mod[:plan.n_latents],
                                     self.w.final_w, self.w.final_b, self._pred_ptrs, self.peers.world,
                                     plan.lat_h, plan.lat_w, norm_weight=self.w.norm, rms_eps=self.eps)
            yield "pred"
            if self.euler_mode is not None:
                # every rank applies the same update to its full copy of the latents (64 KB): latents never travel
                use_cfg, x1_mode = self.euler_mode
                ops.cfg_euler(self.z, self.pred, use_cfg, x1_mode, scalars_dev=self.scalars, vel_out=self.vel[:n_half])

    def predict_steps(self):
        """Generator form of ``predict`` without CUDA-graph capture (lockstep execution)."""
        if not self.prefilled:
            raise RuntimeError("prefill() must run before predict()")
        yield from self._predict_kernels()

    def predict(self) -> torch.Tensor:
        """One denoising forward of the active rows.  Inputs: ``self.z`` (latents), ``self.t``
        (one timestep per latent); output ``self.pred`` (raw model prediction per latent; in a
        sequence-parallel group the complete prediction, on every rank)."""
        if not self.prefilled:
            raise RuntimeError("prefill() must run before predict()")
        if self.peers is not None and self.peers.lockstep:
            raise RuntimeError("virtual ranks (LocalPeerGroup) run through engine.run_lockstep")
        if not self.use_cuda_graph:
            self._drive(self._predict_kernels())
            return self.pred
        key = (self.uniform_t, self.euler_mode)
        if self._graph is None or self._graph_key != key:
            self._graph_key = key
            if self.euler_mode is not None:        # the warm-up and the capture pass must not advance the latents
                z_keep = self.z.clone()
            self._drive(self._predict_kernels())   # warm-up (sets function attributes, fills caches)
            torch.cuda.synchronize()
            if self.peers is not None:
                self.peers.host_barrier()          # nobody spins on a peer while anybody captures
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                self._drive(self._predict_kernels())
            self._graph = g
            if self.euler_mode is not None:
                self.z.copy_(z_keep)
            if self.peers is not None:
                self.peers.host_barrier()
        self._graph.replay()
        return self.pred

    @property
    def launches_per_predict(self) -> int:
        """Kernels of this library per predict() (torch's two replicate copies under uniform_t are
        not counted)."""
        sync = 0
        if self.peers is not None and not self.peers.lockstep:
            sync = 2 if self.plan is not None and self.plan.partition == "sequences" else self.L + 1
        return 6 + 1 + 8 * self.L + 1 + sync

    @property
    def launches_per_prefill(self) -> int:
        return (1 + 8 * (self.L - 1) + 3) if self.plan.prefix.rows else 0


def run_lockstep(generators):
    """Advance the kernel-sequence generators of several virtual ranks (``LocalPeerGroup``) phase by
    phase on the current stream: everything every rank does before sync point k is enqueued before
    anything any rank does after it, which is what the cross-GPU barrier guarantees between real
    ranks."""
    live = list(generators)
    while live:
        nxt = []
        for g in live:
            try:
                next(g)
                nxt.append(g)
            except StopIteration:
                pass
        if nxt and len(nxt) != len(live):
            raise RuntimeError("virtual ranks disagree on the number of sync points")
        live = nxt
