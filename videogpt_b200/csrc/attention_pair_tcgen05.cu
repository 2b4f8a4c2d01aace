// Clip-block-causal flash attention, production kernel: TWO 128-row query tiles per CTA that
// ping-pong on the tensor pipe (tcgen05 + TMEM + TMA).
//
// Same contract as attention.cu, the mma.sync cross-check kernel (codes instead of a mask, paged KV pools,
// tile classification from per-tile code min/max).  What changed against the first tcgen05 version
// (one query tile, 8 softmax warps, 101 us per launch at cfg2) and why:
//
//   * One CTA owns query tiles A and B of the same (sequence, head): every K/V tile is loaded once
//     for 256 queries (half the L2 -> SMEM traffic), and while the softmax warps of A work on
//     S_A(j) the tensor pipe runs O_B += P_B V(j-1) and S_B(j) -- the MMA issue order is
//     S_A S_B | PV_A S_A' | PV_B S_B' | ...  so neither pipe waits for the other.
//   * One softmax THREAD per query row (4 warps per tile, TMEM lane == row): the row maximum and
//     row sum need no shuffle, no shared memory and no named barrier.
//   * P overwrites S in tensor memory (bf16 pairs in the first 64 columns of the S region) and is
//     the A operand of O += P V; S(j+1) is issued behind P V(j) on the in-order tensor pipe, so
//     the aliasing needs no extra barrier.  TMEM: S_A | S_B | O_A | O_B = 2 x 128 + 2 x D <= 512.
//   * Lazy rescale: the running maximum only moves when a row's tile maximum exceeds it by more
//     than 2^8 (in the exp2 domain), so after the first tiles O is almost never rescaled; the
//     rare rescale is done by the row's own thread (tcgen05.ld / st), warp-uniformly, and rows
//     whose maximum did not move use alpha == 1 exactly -- results do not depend on which rows
//     share a tile (sequence parallelism relies on it).
//   * The code predicate is evaluated only by warps that contain a row that cannot see the whole
//     tile (the two tag rows at a frame start) and on the ragged last tile -- in registers, from key codes
//     prefetched before the wait for S; on a tile whose keys all carry one code such a row simply contributes P = 0.
//
// Warp roles (384 threads = 3 warpgroups): warps 0..3 = softmax / epilogue of tile A, warps 4..7 =
// of tile B (TMEM lane quadrant = warp_idx % 4), warp 8 = TMA producer, warp 9 = TMEM allocator +
// MMA issuer, warp 10 = tile-table builder, warp 11 idle.  Registers are re-balanced per
// warpgroup with setmaxnreg (softmax 224, the rest 56): a softmax thread holds a whole 128-key
// row of S in registers.
#include "common.cuh"
#include "vgpt_internal.h"

#include <cuda.h>

#include <cstdlib>

namespace vgpt {

constexpr int kPairBM = 128;         // queries per tile (two tiles per CTA)
constexpr int kPairBN = 128;         // keys per tile = one KV page
constexpr int kPairThreads = 384;   // 3 warpgroups: softmax A, softmax B, {TMA, MMA, table, idle}

struct AttnSeqP { int32_t q_row0, n_q, kv_len, reserved; };

template <int D>
struct PairCfg {
  static constexpr int kCW = (D == 96) ? 32 : 64;            // elements per swizzled chunk row
  static constexpr int kRowBytes = kCW * 2;                   // 64 (SW64) or 128 (SW128)
  static constexpr uint32_t kLayout = (D == 96) ? kLayoutSW64 : kLayoutSW128;
  static constexpr int kChunks = D / kCW;
  static constexpr int kChunkBytes = 128 * kRowBytes;
  static constexpr int kTileBytes = kChunks * kChunkBytes;    // Q, K or V tile = 128 * D * 2
  static constexpr int kStages = (D == 128) ? 2 : (D == 96 ? 3 : 4);
  static constexpr int kCodeScratch = 8 * 128 * 4;            // one tile of key codes per softmax warp
  static constexpr int kSmem = kTileBytes * (2 + 2 * kStages) + 256 + 3 * 1024 * 4 + kCodeScratch + 1024;   // tiles + barriers + tile table + code scratch + align
  static constexpr int kSmemTrace = kSmem + 10 * 192 * 8;      // + the diagnostic instantiation's stamp rings
  static constexpr int kTmemO = 256;                          // O_A at 256, O_B at 256 + D
  // 64 spare TMEM columns (head_dim <= 96): P gets its own buffer, shared by the two tiles, and
  // S_x(j+1) is issued as soon as the softmax warps have read S_x(j) -- it runs under softmax(j).
  // Otherwise (head_dim 128) P overwrites S and S_x(j+1) follows P V_x(j) on the tensor pipe.
  static constexpr bool kEarlyS = 256 + 2 * D + 64 <= 512;
  static constexpr int kTmemP = 256 + 2 * D;
};

__device__ __forceinline__ float ex2_ftz(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// packed fp32x2 math (Blackwell FFMA2 / FADD2): halves the FMA-pipe instruction count of the
// exponent arguments and of the row sums
__device__ __forceinline__ void ffma2(float& d0, float& d1, float a0, float a1, float b, float c) {
  asm("{\n\t.reg .b64 ra, rb, rc, rd;\n\t"
      "mov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %4};\n\tmov.b64 rc, {%5, %5};\n\t"
      "fma.rn.f32x2 rd, ra, rb, rc;\n\t"
      "mov.b64 {%0, %1}, rd;\n\t}"
      : "=f"(d0), "=f"(d1)
      : "f"(a0), "f"(a1), "f"(b), "f"(c));
}
__device__ __forceinline__ void fadd2(float& d0, float& d1, float a0, float a1) {     // d += a
  asm("{\n\t.reg .b64 ra, rd;\n\t"
      "mov.b64 ra, {%2, %3};\n\tmov.b64 rd, {%0, %1};\n\t"
      "add.rn.f32x2 rd, rd, ra;\n\t"
      "mov.b64 {%0, %1}, rd;\n\t}"
      : "+f"(d0), "+f"(d1)
      : "f"(a0), "f"(a1));
}

// Diagnostic instantiation (VGPT_ATTN_VARIANT=8, tools/attn_trace.py): lane 0 of ten warps of CTA (0, 0, 0) -- the eight
// softmax warps, the TMA producer and the MMA issuer -- stamp the SM clock at every hand-over between the
// roles into a per-warp ring in SHARED memory (a clock read and one st.shared: tens of cycles; a first version stamped
// through a global atomic counter and cost ~900 cycles per event, which distorted the very timeline it recorded:
// profiles/r02b_attn_trace_atomic_stamps.txt); the rings are copied to global memory when the CTA is done.  The default
// instantiation contains none of it.  (The opt-in variants of round 1 -- ragged last KV tile trimmed to ceil16(tail),
// every fourth / every second exponential as a polynomial on the FMA pipe -- were validated and timed in round 2:
// bit-exact / within 2^-7, and 0.3 - 2.8 us SLOWER than the default at cfg2 (profiles/r02b_attn_variants.txt): the MUFU
// unit is 39 % busy, not the limit.  They were deleted.)
constexpr int kTraceRoles = 10, kTracePerRole = 192;     // warps 0..7 (softmax A / B, every quadrant), 8 (TMA), 9 (MMA)
constexpr int kTraceMax = kTraceRoles * kTracePerRole;
__device__ unsigned long long g_attn_trace[2 * kTraceMax];     // (clock, warp << 40 | tile << 32 | j << 8 | event)
__device__ unsigned int g_attn_trace_n;

enum : int {   // trace events
  kEvSFree = 1, kEvSIssued = 2, kEvPFull = 3, kEvPVIssued = 4, kEvMmaTop = 5, kEvMmaFenced = 6,   // MMA warp
  kEvSFull = 10, kEvSRead = 11, kEvExpDone = 12, kEvPBufFree = 13, kEvPWritten = 14, kEvEpilogue = 15,   // softmax warps
  kEvKvEmpty = 20, kEvKvIssued = 21,                                                // TMA warp
  kEvStart = 30, kEvTableDone = 31, kEvEnd = 32
};

template <bool TRACE>
struct Tracer {
  uint2* ring = nullptr;       // this warp's ring in shared memory (null: this thread does not record)
  int n = 0;
  __device__ __forceinline__ void operator()(int tile, int j, int ev) {
    if constexpr (TRACE) {
      if (ring != nullptr && n < kTracePerRole) {
        ring[n++] = make_uint2((unsigned)clock(), ((unsigned)(tile & 0xff) << 28) | ((unsigned)(j & 0xfffff) << 8) | (unsigned)ev);
      }
    }
  }
};

constexpr float kMasked = -1e30f;        // score of a hidden (query, key) pair on a mixed tile: exp2 gives 0 exactly
constexpr int kTileUniform = 1 << 30;   // flag in the tile table's logical index: all 128 keys of the tile share one code
constexpr int kPairMaxTiles = 1024;      // KV tiles per sequence the per-CTA tile table can hold (128 K tokens)

template <int D, bool TRACE>
__global__ void __launch_bounds__(kPairThreads, 1)
attn_pair_tcgen05_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_k,
                         const __grid_constant__ CUtensorMap tmap_v, __nv_bfloat16* __restrict__ out, int out_ld,
                         const int32_t* __restrict__ page_table, int max_pages,
                         const AttnSeqP* __restrict__ seqs, const int32_t* __restrict__ q_code,
                         const int32_t* __restrict__ k_code, const int32_t* __restrict__ k_tile_minmax,
                         int max_k_tiles64, int H, float scale_log2, int dbg) {
  using C = PairCfg<D>;
  constexpr int kStages = C::kStages;
  extern __shared__ uint8_t smem_raw[];
  __shared__ int s_qmin, s_qmax, s_nvis;
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gen = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t s_q = base;                                   // Q_A, Q_B
  const uint32_t s_kv = s_q + 2 * C::kTileBytes;               // stage s: K at s_kv + 2*s*tile, V right after
  const uint32_t bars = s_kv + 2 * kStages * C::kTileBytes;
  const uint32_t bar_q = bars;
  auto bar_kv_full = [&](int s) { return bars + 8u * (1 + s); };
  auto bar_kv_empty = [&](int s) { return bars + 8u * (1 + kStages + s); };
  auto bar_s_full = [&](int x) { return bars + 8u * (1 + 2 * kStages + x); };
  auto bar_p_full = [&](int x) { return bars + 8u * (3 + 2 * kStages + x); };
  auto bar_o_full = [&](int x) { return bars + 8u * (5 + 2 * kStages + x); };
  auto bar_s_free = [&](int x) { return bars + 8u * (7 + 2 * kStages + x); };
  const uint32_t tmem_slot = bars + 8u * (9 + 2 * kStages);
  uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(gen + (tmem_slot - base));
  // per-CTA table of the KV tiles some query of this CTA can see (built once; every role walks it
  // from shared memory instead of re-deriving it from global memory inside its loop)
  int32_t* t_page = reinterpret_cast<int32_t*>(gen + (bars - base) + 256);          // pool page of tile i
  int32_t* t_tmax = t_page + kPairMaxTiles;                                         // max key code of tile i
  int32_t* t_kt = t_tmax + kPairMaxTiles;                                           // logical tile index

  // Grid (head, query pair, sequence), head fastest.  CTAs are dispatched in linear block order to
  // whichever SM frees up first, and a CTA's work is (query tiles) x (KV tiles it can see): about 2.2 CTAs of very
  // different sizes per SM at cfg2, so the order decides the makespan.  With the heads innermost all full pairs of
  // the longest (first: conditional) sequence start before its half-empty last pair and before the short
  // unconditional sequence -- longest first.  (With the pair index innermost the first wave mixed 4 full pairs :
  // 1 tail per head and left 12 full-size CTAs for a second round.  Sending the half-empty last pair of every
  // sequence to the very end instead -- list scheduling of the measured CTA durations at cfg2 predicts makespan 109
  // against 114 thousand cycles -- measured no gain at cfg2 (66.9 vs 66.6 us) and a loss at cfg3, where the tail of
  // the long sequence walks 73 KV tiles and must not start last: 208 vs 183 us; profiles/r02m_attn_bench_tails_last.txt.)
  const int seq_id = blockIdx.z, head = blockIdx.x;
  const AttnSeqP sq = seqs[seq_id];
  const int q0 = blockIdx.y * 2 * kPairBM;
  if (q0 >= sq.n_q) return;
  const int rows_cta = min(2 * kPairBM, sq.n_q - q0);          // valid rows of A and B together
  const bool has_b = rows_cta > kPairBM;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  Tracer<TRACE> tr;
  int4* t_kcode = reinterpret_cast<int4*>(t_kt + kPairMaxTiles);                    // [softmax warp][32 lanes] key codes of a tile
  [[maybe_unused]] uint2* trace_rings = reinterpret_cast<uint2*>(t_kcode + 8 * 32);
  __shared__ int s_trace_n[kTraceRoles];
  if constexpr (TRACE) {
    if ((blockIdx.x | blockIdx.y | blockIdx.z) == 0 && lane == 0 && warp < kTraceRoles) tr.ring = trace_rings + warp * kTracePerRole;
  }

  // ---- CTA-wide max of the valid query codes (tile classification) -----------------------------
  if (threadIdx.x == 0) { s_qmin = 0x7fffffff; s_qmax = (int)0x80000000; s_nvis = 0; }
  __syncthreads();
  if ((int)threadIdx.x < rows_cta && threadIdx.x < 2 * kPairBM) atomicMax(&s_qmax, q_code[sq.q_row0 + q0 + threadIdx.x]);
  if (warp == 8 && lane == 0) {
    tma_prefetch_desc(&tmap_q); tma_prefetch_desc(&tmap_k); tma_prefetch_desc(&tmap_v);
    mbar_init(bar_q, 1);
    for (int s = 0; s < kStages; ++s) { mbar_init(bar_kv_full(s), 1); mbar_init(bar_kv_empty(s), 1); }
    for (int x = 0; x < 2; ++x) {
      mbar_init(bar_s_full(x), 1); mbar_init(bar_p_full(x), 4); mbar_init(bar_o_full(x), 1); mbar_init(bar_s_free(x), 4);
    }
    fence_barrier_init();
  }
  if (warp == 9) tmem_alloc<512>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot_ptr;
  const int n_kt = (sq.kv_len + kPairBN - 1) / kPairBN;
  tr(0, 0, kEvStart);
  if (warp == 8) {                        // Q does not need the table: start its load now
    if (elect_one_sync()) {
      mbar_arrive_expect_tx(bar_q, (has_b ? 2 : 1) * C::kTileBytes);
      for (int x = 0; x < (has_b ? 2 : 1); ++x)
#pragma unroll
        for (int c = 0; c < C::kChunks; ++c)
          tma_load_2d(s_q + x * C::kTileBytes + c * C::kChunkBytes, &tmap_q, bar_q, head * D + c * C::kCW,
                      sq.q_row0 + q0 + x * kPairBM);
    }
    __syncwarp();
  } else if (warp == 10) {                // ordered compaction of the visible tiles, 32 per step
    const int q_max = s_qmax;
    const int32_t* mm = k_tile_minmax + (size_t)seq_id * max_k_tiles64 * 2;   // (min, max) per 64 keys
    const int32_t* pt = page_table + (size_t)seq_id * max_pages;
    int n = 0;
    for (int b0 = 0; b0 < n_kt; b0 += 32) {
      const int kt = b0 + lane;
      bool vis = false;
      int tmax = 0, page = 0, uni = 0;
      if (kt < n_kt) {
        const int4 m4 = *reinterpret_cast<const int4*>(mm + 4 * kt);
        vis = min(m4.x, m4.z) <= q_max;               // else: fully masked for every query of this CTA
        tmax = max(m4.y, m4.w);
        uni = min(m4.x, m4.z) == tmax ? kTileUniform : 0;     // every key of the tile carries the same code
        page = pt[kt];
      }
      const uint32_t bal = __ballot_sync(0xffffffffu, vis);
      if (vis) {
        const int i = n + __popc(bal & ((1u << lane) - 1u));
        t_page[i] = page; t_tmax[i] = tmax; t_kt[i] = kt | uni;
      }
      n += __popc(bal);
    }
    if (lane == 0) s_nvis = n;
  }
  __syncthreads();
  const int n_vis = s_nvis;
  tr(0, n_vis, kEvTableDone);

  if (warp >= 8) {
    // 256 x 224 + 128 x 56 = 64512 = the 384 x 168 registers the CTA was launched with.  (A first version gave the
    // variants 64 here: 65536 > 64512, and setmaxnreg.inc of the softmax warps waited forever.)
    asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");
  if (warp == 8) {
    // =================================== TMA producer ===================================
    // (whole warp runs the loop; one elected lane issues)
    int stage = 0; uint32_t phase = 0;
    for (int i = 0; i < n_vis; ++i) {
      const int row = (t_page[i] * H + head) * kPairBN;         // pool viewed as [(page*H + head)*128 + tok][D]
      mbar_wait_relaxed(bar_kv_empty(stage), phase ^ 1);
      tr(0, i, kEvKvEmpty);
      if ((dbg & 8) && i >= kStages) {                          // timing probe: operands not refreshed
        if (elect_one_sync()) mbar_arrive(bar_kv_full(stage));
      } else if (elect_one_sync()) {
        const uint32_t sk = s_kv + 2 * stage * C::kTileBytes, sv = sk + C::kTileBytes;
        mbar_arrive_expect_tx(bar_kv_full(stage), 2 * C::kTileBytes);
#pragma unroll
        for (int c = 0; c < C::kChunks; ++c) tma_load_2d(sk + c * C::kChunkBytes, &tmap_k, bar_kv_full(stage), c * C::kCW, row);
#pragma unroll
        for (int c = 0; c < C::kChunks; ++c) tma_load_2d(sv + c * C::kChunkBytes, &tmap_v, bar_kv_full(stage), c * C::kCW, row);
      }
      __syncwarp();
      tr(0, i, kEvKvIssued);
      if (++stage == kStages) { stage = 0; phase ^= 1; }
    }
  } else if (warp == 9) {
    // =================================== MMA issuer ===================================
    // Warp-uniform control flow; the MMAs / commits are issued by one elected lane.
    // (One issuing warp per query tile -- warp 10 for tile B -- was tried once the trace showed this warp's serial chain
    // of loop code, ~150-cycle barrier polls and MMA-queue blocking at 1440 cycles per tile: bit-identical, and no faster
    // (cfg2 66.6 -> 68.7 us, cfg5 613 -> 600 us, profiles/r02i_attn_bench_two_issuers.txt): with the issuers side by side
    // the period stayed at 2900 cycles because the softmax warps' own chain per tile -- S read 175, exponentials
    // 1600-1900 with the other tile's warp on the same MUFU unit, shared P buffer 230-430, P write 160, next barrier 400 --
    // is that long.  Deleted.)
    constexpr uint32_t idesc_s = make_idesc_bf16(128, kPairBN);        // S = Q K^T (both K-major)
    constexpr uint32_t idesc_o = make_idesc_bf16(128, D, 0, 1);        // O = P V   (V MN-major)
    // Descriptors: everything but the 14-bit start address is constant, and the start addresses of one issue differ by
    // compile-time offsets -- one add per descriptor on the low word.  (Rebuilding every descriptor from its byte
    // address cost ~70 uniform-datapath instructions in front of each group of MMAs, four times per KV tile, on a warp
    // that shares its scheduler with two softmax warps: the trace showed 300 - 700 idle cycles of the MMA warp between
    // hand-overs it was not waiting for, profiles/r02d_attn_trace.txt.)
    constexpr uint64_t kDescQK = make_smem_desc(0, 16, 8 * C::kRowBytes, C::kLayout);                // K-major Q / K
    constexpr uint64_t kDescV = make_smem_desc(0, C::kChunkBytes, 8 * C::kRowBytes, C::kLayout);     // MN-major V
    constexpr uint32_t kStageStep = (2 * C::kTileBytes) >> 4;
    auto desc = [](uint64_t fixed, uint32_t lo) { return (fixed & 0xffffffff00000000ull) | (uint64_t)lo; };
    const uint32_t q_lo0 = (uint32_t)kDescQK | ((s_q >> 4) & 0x3fffu);
    const uint32_t q_lo1 = q_lo0 + (C::kTileBytes >> 4);
    const uint32_t k_lo0 = (uint32_t)kDescQK | ((s_kv >> 4) & 0x3fffu);
    const uint32_t v_lo0 = (uint32_t)kDescV | (((s_kv + C::kTileBytes) >> 4) & 0x3fffu);
    auto issue_s = [&](int x, int stage) {
      if (elect_one_sync()) {
        const uint32_t ql = x ? q_lo1 : q_lo0, kl = k_lo0 + (uint32_t)stage * kStageStep;
        const uint32_t td = tmem + x * 128;
#pragma unroll
        for (int c = 0; c < C::kChunks; ++c) {
#pragma unroll
          for (int ks = 0; ks < C::kCW / 16; ++ks) {
            const uint32_t off = (uint32_t)(c * C::kChunkBytes + ks * 32) >> 4;
            umma_f16_ss(td, desc(kDescQK, ql + off), desc(kDescQK, kl + off), idesc_s, (c | ks) ? 1u : 0u);
          }
        }
        if (!(dbg & 32)) umma_commit(bar_s_full(x));          // (dbg & 32: timing probe, fewer commits)
      }
      __syncwarp();
    };
    auto issue_pv = [&](int x, int stage, int j, bool release_kv) {
      if (elect_one_sync()) {
        const uint32_t vl = v_lo0 + (uint32_t)stage * kStageStep;
        const uint32_t td = tmem + C::kTmemO + x * D, tp = tmem + (C::kEarlyS ? C::kTmemP : x * 128);
#pragma unroll
        for (int ks = 0; ks < kPairBN / 16; ++ks)
          umma_f16_ts(td, tp + ks * 8, desc(kDescV, vl + (uint32_t)((ks * 16 * C::kRowBytes) >> 4)), idesc_o,
                      (j > 0 || ks > 0) ? 1u : 0u);
        if (!(dbg & 32) || j == n_vis - 1) umma_commit(bar_o_full(x));
        if (release_kv) umma_commit(bar_kv_empty(stage));     // K(j), V(j) free once everything issued so far is done
      }
      __syncwarp();
    };
    mbar_wait(bar_q, 0);
    int stage_s = 0; uint32_t phase_s = 0;
    int stage_o = 0;
    if (n_vis > 0) {                            // prologue: S_A(0), S_B(0)
      mbar_wait(bar_kv_full(stage_s), phase_s);
      tc_fence_after();
      issue_s(0, stage_s);
      if (has_b) issue_s(1, stage_s);
      if (++stage_s == kStages) { stage_s = 0; phase_s ^= 1; }
    }
    if constexpr (C::kEarlyS) {
      // S_x(j+1) as soon as S_x(j) sits in the softmax warps' registers (s_free), P V_x(j) when
      // P_x(j) is in the shared P buffer (p_full); fixed order A, B -- the two tiles fall into a
      // half-period stagger.  (An event loop polling all four barriers was measured slower:
      // it competes with the softmax warps for issue slots.  Issuing every barrier poll early -- a non-blocking test
      // in front of the MMA group / P hand-over that precedes the wait, in this warp and in the softmax warps -- was
      // measured too: 67.8 / 187.9 / 622.4 us at cfg2 / cfg3 / cfg5 against 66.6 / 183.3 / 613.4,
      // profiles/r02n_attn_bench_early_polls.txt.  Deleted.)
      for (int j = 0; j < n_vis; ++j) {
        const bool has_next = j + 1 < n_vis;
        for (int x = 0; x < (has_b ? 2 : 1); ++x) {
          if (has_next) {
            tr(x, j + 1, kEvMmaTop);
            if (!(dbg & 4)) mbar_wait(bar_s_free(x), j & 1);
            if (x == 0) mbar_wait(bar_kv_full(stage_s), phase_s);
            tr(x, j + 1, kEvSFree);
            tc_fence_after();
            tr(x, j + 1, kEvMmaFenced);
            issue_s(x, stage_s);
            tr(x, j + 1, kEvSIssued);
          }
          if (!(dbg & 4)) mbar_wait(bar_p_full(x), j & 1);
          tr(x, j, kEvPFull);
          tc_fence_after();
          issue_pv(x, stage_o, j, x == (has_b ? 1 : 0));
          tr(x, j, kEvPVIssued);
        }
        if (has_next && ++stage_s == kStages) { stage_s = 0; phase_s ^= 1; }
        if (++stage_o == kStages) stage_o = 0;
      }
    } else {
      for (int j = 0; j < n_vis; ++j) {
        const bool has_next = j + 1 < n_vis;
        mbar_wait(bar_p_full(0), j & 1);          // P_A(j) in TMEM (and O_A rescaled if it had to be)
        tc_fence_after();
        issue_pv(0, stage_o, j, !has_b);
        if (has_next) {
          mbar_wait(bar_kv_full(stage_s), phase_s);
          tc_fence_after();
          issue_s(0, stage_s);    // S_A(j+1) overwrites P_A(j): behind P V_A(j) in pipe order
        }
        if (has_b) {
          mbar_wait(bar_p_full(1), j & 1);
          tc_fence_after();
          issue_pv(1, stage_o, j, true);
          if (has_next) issue_s(1, stage_s);
        }
        if (has_next && ++stage_s == kStages) { stage_s = 0; phase_s ^= 1; }
        if (++stage_o == kStages) stage_o = 0;
      }
    }
  }
  } else {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 224;");
    // ============================ softmax / rescale / epilogue ============================
    // (Two warps per row block -- 64 key columns each, the row maximum exchanged through shared memory and a named
    // barrier of their 64 threads, partial row sums added in a fixed order, 640 threads -- were built on the theory that
    // one warp's pass is bound by its own issue occupancy (128 MUFU x 8 cycles + ~235 packed FMA / pack / max x 2 = 1500
    // cycles, what the trace shows) and that a second warp on the same scheduler would hide one's MUFU cycles under the
    // other's FMA-pipe cycles.  Bit-identical in every test; 79.3 us at cfg2 with the single MMA issuer, 68.4 us with one
    // issuer per query tile, against 66.4 us for this layout (cfg5: 765 / 632 / 606 us) -- 104 registers per thread spill
    // inside the loop and the exchange adds a barrier to every tile.  profiles/r02p_attn_bench_8_softmax_warps.txt.  Deleted.)
    const int x = warp >> 2;                                    // 0 = tile A, 1 = tile B
    if (x == 0 || has_b) {
      const int quad = warp & 3;
      const int row = quad * 32 + lane;                         // row of the tile == TMEM lane
      const int rows_here = min(kPairBM, rows_cta - x * kPairBM);
      const uint32_t lane_addr = (uint32_t)(quad * 32) << 16;
      const uint32_t t_s = tmem + lane_addr + x * 128;          // S_x (P_x = its first 64 columns)
      const uint32_t t_o = tmem + lane_addr + C::kTmemO + x * D;
      const uint32_t t_p = tmem + lane_addr + C::kTmemP;        // shared P buffer (kEarlyS)
      const int grow = sq.q_row0 + q0 + x * kPairBM + row;
      const bool valid = row < rows_here;
      const int qc = valid ? q_code[grow] : 0x7fffffff;         // padding rows: see everything, never stored
      const int32_t* kc = k_code + (size_t)seq_id * max_pages * kPairBN;
      const float thresh = 8.0f / scale_log2;                   // lazy rescale: 2^8 head-room
      float m_run = -INFINITY, l_run = 0.f;
      for (int j = 0; j < n_vis; ++j) {
        if (dbg & 4) continue;                                    // timing probe: free-running tensor pipe
        const int tmax = t_tmax[j], ktu = t_kt[j];                // (shared memory, before the wait)
        const int kt = ktu & (kTileUniform - 1);
        // Rows that cannot see the whole tile (the two tag rows at a frame start; everybody on the ragged last tile):
        //   * tile with ONE key code (most: 256 patch tokens per frame): such a row sees none of it -- it runs the
        //     same instructions with multiplier 0 / offset -inf (P = 0 exactly) and leaves the row maximum alone;
        //   * mixed tile: the 128 key codes are fetched BEFORE the wait for S (one int4 per lane: the latency hides
        //     under the MMAs), staged in shared memory and applied to S in registers with two FMA-pipe instructions
        //     per score.
        // (First version: S patched in place in tensor memory, 32 columns at a time, codes loaded inside the loop:
        // 2200 cycles per tile for the whole CTA to protect two rows -- profiles/r02d_attn_trace.txt, tiles 8..16.)
        const bool ragged = (kt + 1) * kPairBN > sq.kv_len;
        const bool hidden = qc < tmax;
        const bool some = ragged || __any_sync(0xffffffffu, hidden);
        const bool elementwise = some && (ragged || !(ktu & kTileUniform));     // warp-uniform
        const bool blind = some && !elementwise && hidden;                        // per row
        int4 kc4 = make_int4(0, 0, 0, 0);
        if (elementwise) kc4 = __ldg(reinterpret_cast<const int4*>(kc + kt * kPairBN) + lane);
        mbar_wait(bar_s_full(x), j & 1);
        tr(x, j, kEvSFull);
        tc_fence_after();
        if (dbg & 1) {                                            // timing probe: tensor-pipe chain only
          tc_fence_before();
          __syncwarp();
          if (lane == 0) { if (C::kEarlyS) mbar_arrive(bar_s_free(x)); mbar_arrive(bar_p_full(x)); }
          continue;
        }
        uint32_t s[128];
#pragma unroll
        for (int c = 0; c < 4; ++c)
          tmem_ld_32x32b_x32(t_s + c * 32, *reinterpret_cast<uint32_t(*)[32]>(&s[c * 32]));
        tmem_ld_wait();
        if constexpr (C::kEarlyS) {                                // S_x is free: S_x(j+1) may be issued now
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(bar_s_free(x));
        }
        tr(x, j, kEvSRead);
        if (elementwise) {
          // hidden(q, k) = code_k > code_q, as fp32 arithmetic on the FMA pipe: sat(code_k - code_q) is exactly 0 or 1
          // (codes are integers < 2^24, or the INT_MAX sentinel of padding / evicted keys), and s - 1e30 * hidden
          // leaves a visible score untouched bit for bit.  (Integer compare + select, or sign-mask + LOP3, run on
          // the half-rate ALU pipe: 256 - 384 instructions x 2 cycles put +1300 cycles on the one warp the whole
          // tile waits for -- profiles/r02g_attn_trace_all_warps.txt, tiles 8 / 10 / 12 / 14.)
          float4* my = reinterpret_cast<float4*>(t_kcode + warp * 32);
          my[lane] = make_float4((float)kc4.x, (float)kc4.y, (float)kc4.z, (float)kc4.w);
          __syncwarp();
          const float nqc = -(float)qc;
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            const float4 c4 = my[i];                              // broadcast read
            s[4 * i + 0] = __float_as_uint(fmaf(__saturatef(c4.x + nqc), kMasked, __uint_as_float(s[4 * i + 0])));
            s[4 * i + 1] = __float_as_uint(fmaf(__saturatef(c4.y + nqc), kMasked, __uint_as_float(s[4 * i + 1])));
            s[4 * i + 2] = __float_as_uint(fmaf(__saturatef(c4.z + nqc), kMasked, __uint_as_float(s[4 * i + 2])));
            s[4 * i + 3] = __float_as_uint(fmaf(__saturatef(c4.w + nqc), kMasked, __uint_as_float(s[4 * i + 3])));
          }
          __syncwarp();                                           // scratch is rewritten for the next tile
        }
        float mx0 = -INFINITY, mx1 = -INFINITY, mx2 = -INFINITY, mx3 = -INFINITY;
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          mx0 = fmaxf(mx0, __uint_as_float(s[4 * i + 0]));
          mx1 = fmaxf(mx1, __uint_as_float(s[4 * i + 1]));
          mx2 = fmaxf(mx2, __uint_as_float(s[4 * i + 2]));
          mx3 = fmaxf(mx3, __uint_as_float(s[4 * i + 3]));
        }
        float mx = fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3));
        if (blind || mx < 0.1f * kMasked) mx = -INFINITY;         // the row sees no key of this tile
        const bool grew = mx > m_run + thresh;                   // also true for the first finite maximum
        const float m_new = grew ? mx : m_run;
        const float sub = (m_new == -INFINITY) ? 0.f : __fmul_rn(m_new, scale_log2);
        // alpha == 1 EXACTLY for rows whose reference maximum did not move
        const float alpha = !grew ? 1.f
                            : (m_run == -INFINITY) ? 0.f
                                                   : ex2_ftz(__fsub_rn(__fmul_rn(m_run, scale_log2), sub));
        float sum0 = 0.f, sum1 = 0.f;
        const float nsub = blind ? -INFINITY : -sub, mul = blind ? 0.f : scale_log2;
#pragma unroll
        for (int i = 0; i < 64; ++i) {
          float p0, p1;
          ffma2(p0, p1, __uint_as_float(s[2 * i]), __uint_as_float(s[2 * i + 1]), mul, nsub);
          if (!(dbg & 2)) { p0 = ex2_ftz(p0); p1 = ex2_ftz(p1); }  // (dbg & 2: timing probe without MUFU)
          fadd2(sum0, sum1, p0, p1);
          s[i] = pack_bf16x2(p0, p1);                             // P overwrites the dead half of s[]
        }
        tr(x, j, kEvExpDone);
        if constexpr (C::kEarlyS) {
          // the P buffer is shared: its previous reader is P V of the other tile (B: tile j of A;
          // A: tile j-1 of B), or of this tile when it is alone
          const int y = has_b ? (x ^ 1) : 0;
          const int need = (x == 1) ? j : j - 1;                   // index of that P V
          if (need >= 0) {
            mbar_wait(bar_o_full(y), need & 1);
            tc_fence_after();
          }
          tr(x, j, kEvPBufFree);
          tmem_st_32x32b_x32(t_p, *reinterpret_cast<uint32_t(*)[32]>(&s[0]));
          tmem_st_32x32b_x32(t_p + 32, *reinterpret_cast<uint32_t(*)[32]>(&s[32]));
        } else {
          tmem_st_32x32b_x32(t_s, *reinterpret_cast<uint32_t(*)[32]>(&s[0]));
          tmem_st_32x32b_x32(t_s + 32, *reinterpret_cast<uint32_t(*)[32]>(&s[32]));
        }
        l_run = l_run * alpha + (sum0 + sum1);
        m_run = m_new;
        if (j > 0 && __any_sync(0xffffffffu, grew)) {
          // rare after the first tiles: O_x(j-1) must be complete, then scale this row
          mbar_wait(bar_o_full(x), (j - 1) & 1);
          tc_fence_after();
#pragma unroll 1
          for (int u = 0; u < D / 32; ++u) {
            uint32_t o[32];
            tmem_ld_32x32b_x32(t_o + u * 32, o);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
            tmem_st_32x32b_x32(t_o + u * 32, o);
          }
        }
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_p_full(x));               // one arrival per warp (4 per tile)
        tr(x, j, kEvPWritten);
      }
      // ---- epilogue: O / l -> bf16 -> global ------------------------------------------------
      if (n_vis > 0) {
        mbar_wait(bar_o_full(x), (n_vis - 1) & 1);
        tc_fence_after();
      }
      tr(x, n_vis, kEvEpilogue);
      const float inv = l_run > 0.f ? 1.f / l_run : 0.f;
      __nv_bfloat16* orow = out + (size_t)grow * out_ld + head * D;
#pragma unroll
      for (int u = 0; u < D / 32; ++u) {
        uint32_t o[32];
        if (n_vis > 0) {
          tmem_ld_32x32b_x32(t_o + u * 32, o);
          tmem_ld_wait();
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i) o[i] = 0u;
        }
        if (valid) {
#pragma unroll
          for (int v4 = 0; v4 < 4; ++v4) {
            uint32_t ww[4];
#pragma unroll
            for (int h2 = 0; h2 < 4; ++h2)
              ww[h2] = pack_bf16x2(__uint_as_float(o[v4 * 8 + h2 * 2]) * inv, __uint_as_float(o[v4 * 8 + h2 * 2 + 1]) * inv);
            reinterpret_cast<uint4*>(orow + u * 32)[v4] = make_uint4(ww[0], ww[1], ww[2], ww[3]);
          }
        }
      }
    }
  }

  tr(0, 0, kEvEnd);
  if constexpr (TRACE) {
    if (tr.ring != nullptr) s_trace_n[(tr.ring - trace_rings) / kTracePerRole] = tr.n;
  }
  tc_fence_before();
  __syncthreads();
  if constexpr (TRACE) {
    if ((blockIdx.x | blockIdx.y | blockIdx.z) == 0) {       // rings -> global memory, role after role
      int base = 0;
      for (int r = 0; r < kTraceRoles; ++r) {
        const int n = s_trace_n[r];
        for (int i = threadIdx.x; i < n; i += kPairThreads) {
          const uint2 e = trace_rings[r * kTracePerRole + i];
          g_attn_trace[2 * (base + i)] = e.x;
          g_attn_trace[2 * (base + i) + 1] = ((unsigned long long)r << 40) | ((unsigned long long)(e.y >> 28) << 32) |
                                             (unsigned long long)(e.y & 0x0fffffffu);
        }
        base += n;
      }
      if (threadIdx.x == 0) g_attn_trace_n = (unsigned)base;
    }
  }
  if (warp == 9) {
    tc_fence_after();
    tmem_dealloc<512>(tmem);
  }
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
static int debug_attn_flags() {     // VGPT_DEBUG_ATTN_FLAGS: timing probes only (results are garbage)
  static const int f = [] { const char* e = getenv("VGPT_DEBUG_ATTN_FLAGS"); return e ? atoi(e) : 0; }();
  return f;
}

static bool attn_trace_on() {      // VGPT_ATTN_VARIANT=8: the diagnostic instantiation (head_dim 96); read at every launch
  const char* e = getenv("VGPT_ATTN_VARIANT");
  return e && atoi(e) == 8;
}

template <int D, bool TRACE>
static int launch_attn_pair(const void* q, int q_ld, int q_rows, void* out, int out_ld, const void* k_pool,
                            const void* v_pool, int total_pages, const int32_t* page_table, int max_pages,
                            const void* seqs, int num_seqs, int q_pairs, const int32_t* q_code,
                            const int32_t* k_code, const int32_t* k_tile_minmax, int max_k_tiles64, int H,
                            float scale, cudaStream_t s) {
  using C = PairCfg<D>;
  const CUtensorMapSwizzle swz = (D == 96) ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B;
  CUtensorMap tq, tk, tv;
  cuuint32_t estr[2] = {1, 1};
  cuuint32_t box[2] = {(cuuint32_t)C::kCW, 128};
  {
    cuuint64_t dims[2] = {(cuuint64_t)q_ld, (cuuint64_t)q_rows};
    cuuint64_t strides[1] = {(cuuint64_t)q_ld * 2};
    int rc = encode_tensor_map(&tq, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(q), dims, strides, box, estr, swz);
    if (rc) return rc;
  }
  {
    cuuint64_t dims[2] = {(cuuint64_t)D, (cuuint64_t)total_pages * H * 128};
    cuuint64_t strides[1] = {(cuuint64_t)D * 2};
    int rc = encode_tensor_map(&tk, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(k_pool), dims, strides, box, estr, swz);
    if (rc) return rc;
    rc = encode_tensor_map(&tv, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(v_pool), dims, strides, box, estr, swz);
    if (rc) return rc;
  }
  auto kern = attn_pair_tcgen05_kernel<D, TRACE>;
  constexpr int smem = TRACE ? C::kSmemTrace : C::kSmem;
  // per launch: the attribute is per device, and a process may drive several devices (cheap, capture-safe)
  VGPT_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  dim3 grid(H, q_pairs, num_seqs);
  kern<<<grid, kPairThreads, smem, s>>>(tq, tk, tv, (__nv_bfloat16*)out, out_ld, page_table, max_pages,
                                            (const AttnSeqP*)seqs, q_code, k_code, k_tile_minmax,
                                            max_k_tiles64, H, scale * 1.4426950408889634f, debug_attn_flags());
  VGPT_CHECK_LAUNCH();
  return 0;
}

// Diagnostic: copy the events recorded by the last launches of the kVarTrace variant (and reset the
// counter).  out: 2 * max_events uint64 (clock, tag) in host or device memory; returns the number of
// events through *n_events.
int attn_trace_read(void* out, int max_events, int* n_events, cudaStream_t s) {
  VGPT_CHECK_ARG(out && n_events && max_events > 0, "vgpt_debug_attn_trace: bad arguments");
  VGPT_CHECK_CUDA(cudaStreamSynchronize(s));
  unsigned int n = 0;
  VGPT_CHECK_CUDA(cudaMemcpyFromSymbol(&n, g_attn_trace_n, sizeof(n)));
  if (n > (unsigned)kTraceMax) n = kTraceMax;
  if (n > (unsigned)max_events) n = max_events;
  if (n) VGPT_CHECK_CUDA(cudaMemcpyFromSymbol(out, g_attn_trace, (size_t)n * 2 * sizeof(unsigned long long)));
  const unsigned int zero = 0;
  VGPT_CHECK_CUDA(cudaMemcpyToSymbol(g_attn_trace_n, &zero, sizeof(zero)));
  *n_events = (int)n;
  return 0;
}

int attn_clip_causal_pair(const void* q, int q_ld, int q_rows, void* out, int out_ld, const void* k_pool,
                          const void* v_pool, int total_pages, const int32_t* page_table, int max_pages,
                          const void* seqs, int num_seqs, int max_q_rows, const int32_t* q_code,
                          const int32_t* k_code, const int32_t* k_tile_minmax, int max_k_tiles, int H, int D,
                          float scale, cudaStream_t s) {
  VGPT_CHECK_ARG(q && out && k_pool && v_pool && page_table && seqs && q_code && k_code && k_tile_minmax,
                 "vgpt_attn_clip_causal: null pointer");
  VGPT_CHECK_ARG(H > 0 && (D == 64 || D == 96 || D == 128), "vgpt_attn_clip_causal: head_dim %d unsupported (64, 96, 128)", D);
  VGPT_CHECK_ARG(q_ld % 8 == 0 && out_ld % 8 == 0 && q_ld >= H * D && out_ld >= H * D && q_rows > 0,
                 "vgpt_attn_clip_causal: bad leading dimensions q_ld=%d out_ld=%d", q_ld, out_ld);
  VGPT_CHECK_ARG(((uintptr_t)q & 15) == 0 && ((uintptr_t)out & 15) == 0 && ((uintptr_t)k_pool & 127) == 0 &&
                     ((uintptr_t)v_pool & 127) == 0,
                 "vgpt_attn_clip_causal: misaligned pointer");
  VGPT_CHECK_ARG(max_pages > 0 && total_pages > 0 && max_k_tiles >= 2 * max_pages,
                 "vgpt_attn_clip_causal: max_k_tiles=%d too small for max_pages=%d", max_k_tiles, max_pages);
  VGPT_CHECK_ARG(scale > 0.f, "vgpt_attn_clip_causal: scale must be positive");
  VGPT_CHECK_ARG(max_pages <= kPairMaxTiles, "vgpt_attn_clip_causal: %d pages per sequence (at most %d)", max_pages, kPairMaxTiles);
  if (num_seqs <= 0 || max_q_rows <= 0) return 0;
  const int q_pairs = (max_q_rows + 2 * kPairBM - 1) / (2 * kPairBM);
  const bool trace = D == 96 && attn_trace_on();
#define VGPT_ATTN_CASE(D_, T_)                                                                              \
  if (D == D_ && trace == T_)                                                                                \
    return launch_attn_pair<D_, T_>(q, q_ld, q_rows, out, out_ld, k_pool, v_pool, total_pages, page_table,  \
                                    max_pages, seqs, num_seqs, q_pairs, q_code, k_code, k_tile_minmax,      \
                                    max_k_tiles, H, scale, s);
  VGPT_ATTN_CASE(64, false)
  VGPT_ATTN_CASE(96, false)
  VGPT_ATTN_CASE(96, true)
  VGPT_ATTN_CASE(128, false)
#undef VGPT_ATTN_CASE
  return -1;
}

}  // namespace vgpt
