// bf16 GEMM  C[M,N] = A[M,K] * W[N,K]^T  for the QKV / O / gate-up / down projections of the Phi-3 block
// (reference call sites: LVM/transform/sdpa_transform.py:39,89 and transformers' Phi3MLP), hand-written for
// sm_100a on CTA PAIRS (tcgen05 cta_group::2):
//
//   * operands staged by TMA (cp.async.bulk.tensor, 128-byte swizzle) through a multi-stage shared-memory ring
//     guarded by full / empty mbarriers; a cluster of two CTAs computes a 256 x BN tile: each CTA loads its own 128
//     rows of A and only HALF of the W tile, one tcgen05.mma.cta_group::2 (M = 256) reads both halves, accumulators
//     stay split across the two SMs' tensor memory.  Per SM the L2 -> SMEM fill drops to (128 + BN/2) x 128 B per
//     k-block (64 B/clk at BN = 256; one CTA per 128 x 256 tile needs 96 B/clk and stalled on it), and the 43-cycle
//     per-instruction overhead of a shared-memory-operand MMA disappears (tools/umma_rate.py: N/2 cycles per MMA);
//   * persistent clusters walk the tile list m-fastest so concurrently running clusters share W tiles through L2;
//     two accumulator stages in TMEM overlap the epilogue of tile i with the MMAs of tile i+1;
//   * epilogues fused into the TMEM read-back: plain store, +residual (in place on the residual stream), SwiGLU over
//     block-interleaved gate / up columns ([gate x 16 | up x 16] per 32 packed weight rows);
//   * TAIL ROWS IN THE K-LOOP.  M = 2064 = 8 x 256 + 16 at cfg2 (two tag rows per generated frame), and every
//     sequence-parallel shard has such a remainder: a ninth, 94 % empty row of 256-row tiles costs 11 % of all MMA
//     rows, a whole extra wave for qkv / gate_up, and makes the rank that owns the remainder 25 % slower than its
//     peers.  Instead the LAST FULL tile row is cut into "special" pieces that also compute the tail rows, with the
//     operands swapped, from the W k-blocks already in shared memory:
//         D_tail^T[piece columns, T] += W_piece[columns, 64] * A_tail[T, 64]^T        (T = 16 or 32)
//     i.e. the W rows in the B slot of the stage are the M = 256 operand of a second, tiny MMA per k-step, the tail
//     rows (T/2 per CTA, loaded behind the W rows in the same slot) its N operand, and the accumulator holds the
//     tail transposed (TMEM lane = output column, TMEM column = tail row) next to the main accumulator.  No extra
//     pass over W, no extra tiles; a special piece is narrower than a regular tile (224 / 192 columns) so that both
//     accumulators fit in the 256 TMEM columns of a stage.  Special pieces are dealt to the least-loaded clusters by
//     list scheduling on the host (`Sched::owner`).  A row gets the same bits whichever path computes it.
//     (A first form -- tail tiles of their own after the main tiles -- was bit-exact too but ran latency-bound:
//     156 cycles of MMA work per stage against a stage round trip of thousands; profiles/r02b_gemm_sweep_fused_tail.txt.)
//
//   * (Tried and deleted, round 2: K split in two over 256 x 256 tiles for o_proj / down_proj with an fp32 hand-over
//     through a workspace.  Bit-stable and M-independent, but 3 waves of half tiles cost exactly the MMA cycles of 2 waves
//     of 192-wide tiles, and the hand-over traffic ate the better main-loop efficiency of the wider tile: o_proj 51.9 vs
//     45.4 us, down_proj 92.3 vs 92.3 us at M = 2064; profiles/r02g_gemm_sweep_splitk.txt.)
//
//   * (Also tried: the residual of chunk c + 1 requested before the wait for chunk c's accumulator, so that the epilogue
//     of a cluster's last tile does not pay a global-load latency per chunk: o_proj 45.3 vs 45.1 us, down_proj inside
//     the run-to-run band -- no measurable gain, reverted.)
//
// Pair protocol (cluster of 2 along M; rank 0 = leader): both CTAs' TMA loads are .cta_group::2 and complete_tx on
// the LEADER's full barrier (its arrive.expect_tx accounts for both halves); the leader's elected thread issues the
// MMAs and commits with .multicast::cluster to the empty / tmem-full barriers of BOTH CTAs; each CTA's epilogue warps
// drain their own TMEM half and arrive on the leader's tmem-empty barrier; cluster barriers fence set-up / tear-down.
// Warp roles (192 threads): warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer, warps 2..5 = epilogue
// (TMEM lane quadrant = warp_idx % 4).
#include "common.cuh"
#include "vgpt_internal.h"

#include <cuda.h>

#include <cstdlib>
#include <cstring>

namespace vgpt {

constexpr int kGBlockM = 128;      // rows per CTA (256 per pair)
constexpr int kGBlockK = 64;       // 64 bf16 = 128 bytes = one swizzle-128B row
constexpr int kGThreads = 192;
constexpr int kAccStride = 256;    // TMEM columns per accumulator stage (2 stages = all 512 columns)
constexpr int kTailCol = 224;      // column of the tail accumulator inside a stage
constexpr int kTailMax = 32;       // at most this many tail rows ride in the k-loop
constexpr int kMaxSpecial = 96;    // special pieces per launch (N = 16384 / 192 = 86)

enum : int { kEpiStore = 0, kEpiResidual = 1, kEpiSwiGLU = 2 };

template <int BN>
struct GemmCfg {
  static constexpr int kABytes = kGBlockM * kGBlockK * 2;                   // 16 KB
  // this CTA's half of the W tile; the B slot of the narrower tiles has room for 16 tail rows behind the W rows of
  // a special piece (BN = 256: a 224-wide piece leaves them inside the 128 rows)
  static constexpr int kBRows = (BN == 256) ? 128 : BN / 2 + kTailMax / 2;
  static constexpr int kBBytes = kBRows * kGBlockK * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kStages = (200 * 1024) / kStageBytes > 8 ? 8 : (200 * 1024) / kStageBytes;
  static constexpr int kBarBytes = 512;
  // the swapped tail MMA reads 128 rows from the B slot whatever its size: slack behind the last stage keeps that read
  // inside the allocation (the rows past the piece's own half only feed accumulator lanes that are never read back)
  static constexpr int kSlackBytes = (128 - kBRows) * kGBlockK * 2;
  static constexpr int kSmemBytes = kStages * kStageBytes + kSlackBytes + kBarBytes + 1024;   // + alignment
  // widest special piece: main + tail accumulators inside one 256-column stage; W rows + tail rows inside the B slot
  static constexpr int kSpecialWidth = (BN == 256) ? kTailCol : BN;
};

// Tile schedule of one launch (by value).  Regular tiles: rows [0, 256 * m_tiles_reg) x all N, 256 x BN, tile t ->
// cluster t % clusters (m fastest).  Special pieces (tail_rows > 0): tile row `m_tiles_reg` (rows sp_row0 ..
// sp_row0 + 255, all valid) cut into n_special pieces of sp_width columns (the last may be narrower), each also
// computing rows [sp_row0 + 256, + tail_rows) of its columns; owner[j] = cluster of piece j.
struct Sched {
  int m_tiles_reg;
  int tail_rows;      // 0: no special pieces
  int T;              // tail rows padded to 16 / 32 (the tail MMA's N)
  int sp_width;
  int n_special;
  uint8_t owner[kMaxSpecial];
};

__device__ __forceinline__ void tmem_ld_32x32b_x16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}

// One 32-column chunk of one accumulator row -> 32 bf16 in global memory (store / +residual).
template <int EPI>
__device__ __forceinline__ void store_chunk(const uint32_t (&acc)[32], __nv_bfloat16* __restrict__ out,
                                            const __nv_bfloat16* __restrict__ res) {
  uint4 r[4];
  if constexpr (EPI == kEpiResidual) {
    const uint4* rp = reinterpret_cast<const uint4*>(res);
#pragma unroll
    for (int i = 0; i < 4; ++i) r[i] = rp[i];
  }
  uint4* op = reinterpret_cast<uint4*>(out);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    uint32_t w[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float a = __uint_as_float(acc[i * 8 + j * 2]);
      float b = __uint_as_float(acc[i * 8 + j * 2 + 1]);
      if constexpr (EPI == kEpiResidual) {
        // reference: o_proj / down_proj output is rounded to bf16, then added to the bf16 residual stream and
        // rounded again (Phi3DecoderLayer.forward, transformers 4.47.1)
        const uint32_t rv = (&r[i].x)[j];
        a = rbf(a) + bf16lo(rv);
        b = rbf(b) + bf16hi(rv);
      }
      w[j] = pack_bf16x2(a, b);
    }
    op[i] = make_uint4(w[0], w[1], w[2], w[3]);
  }
}

// SwiGLU: a 32-column chunk holds [gate x 16 | up x 16] of 16 consecutive outputs -> 16 bf16 (Phi3MLP: gate_up rounded
// to bf16, silu rounded, product rounded).
__device__ __forceinline__ void store_chunk_swiglu(const uint32_t (&acc)[32], __nv_bfloat16* __restrict__ out) {
  uint32_t w[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const float g0 = rbf(__uint_as_float(acc[2 * i])), g1 = rbf(__uint_as_float(acc[2 * i + 1]));
    const float u0 = rbf(__uint_as_float(acc[16 + 2 * i])), u1 = rbf(__uint_as_float(acc[16 + 2 * i + 1]));
    w[i] = pack_bf16x2(u0 * rbf(silu_f(g0)), u1 * rbf(silu_f(g1)));
  }
  uint4* op = reinterpret_cast<uint4*>(out);
  op[0] = make_uint4(w[0], w[1], w[2], w[3]);
  op[1] = make_uint4(w[4], w[5], w[6], w[7]);
}

template <int BN, int EPI>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kGThreads, 1)
gemm_bf16_pair_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                      const __grid_constant__ CUtensorMap tmap_wsp, const __grid_constant__ CUtensorMap tmap_t,
                      __nv_bfloat16* __restrict__ C, const __nv_bfloat16* __restrict__ R, int M, int N, int K, int ldc,
                      int flags, const __grid_constant__ Sched sched) {
  using Cfg = GemmCfg<BN>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;        // swizzle-128B atoms: 1 KB aligned
  const uint32_t bar_base = smem_base + Cfg::kStages * Cfg::kStageBytes + Cfg::kSlackBytes;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (Cfg::kStages + s); };
  auto tmem_full_bar = [&](int s) { return bar_base + 8u * (2 * Cfg::kStages + s); };
  auto tmem_empty_bar = [&](int s) { return bar_base + 8u * (2 * Cfg::kStages + 2 + s); };
  const uint32_t tmem_slot = bar_base + 8u * (2 * Cfg::kStages + 4);
  uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int cluster_id = blockIdx.x >> 1, num_clusters = gridDim.x >> 1;
  const int m_tiles = sched.m_tiles_reg;
  const int n_tiles = (N + BN - 1) / BN;
  const int num_tiles = m_tiles * n_tiles;
  const int k_blocks = K / kGBlockK;
  const int n_special = sched.tail_rows > 0 ? sched.n_special : 0;
  const int sp_row0 = m_tiles * 2 * kGBlockM;                     // first row of the special tile row
  const uint32_t tail_off = (uint32_t)(sched.sp_width / 2) * kGBlockK * 2;   // tail rows behind the W rows in the B slot

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_b);
    if (n_special) { tma_prefetch_desc(&tmap_wsp); tma_prefetch_desc(&tmap_t); }
    for (int s = 0; s < Cfg::kStages; ++s) {
      mbar_init(full_bar(s), 1);          // leader producer's arrive.expect_tx covers both CTAs' bytes
      mbar_init(empty_bar(s), 1);         // leader's multicast commit
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(tmem_full_bar(s), 1);     // leader's multicast commit
      mbar_init(tmem_empty_bar(s), 8);    // 4 epilogue warps x 2 CTAs (used in the leader only)
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc_pair<2 * kAccStride>(tmem_slot);
  tc_fence_before();
  cluster_sync_all();                     // barriers of both CTAs initialised, TMEM allocated
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  if (warp == 0) {
    // ================================ TMA producer (both CTAs) ================================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = cluster_id; tile < num_tiles; tile += num_clusters) {
        const int m0 = (tile % m_tiles) * 2 * kGBlockM + rank * kGBlockM;
        const int n0 = (tile / m_tiles) * BN + rank * (BN / 2);
        for (int kb = 0; kb < k_blocks; ++kb) {
          mbar_wait(empty_bar(stage), phase ^ 1);
          const uint32_t sa = smem_base + stage * Cfg::kStageBytes;
          const uint32_t sb = sa + Cfg::kABytes;
          if (flags & 1) {                 // debug: MMA-bound ceiling, operands not refreshed
            if (leader) mbar_arrive(full_bar(stage));
          } else {
            // The peer only issues its loads: its bytes are credited to the leader's barrier, whose phase cannot
            // complete before the leader's own arrive.expect_tx (count 1).  (A remote mbarrier.arrive.release.cluster
            // here serialised the peer's loads: profiles/r01c.)
            if (leader) mbar_arrive_expect_tx(full_bar(stage), 2 * (Cfg::kABytes + (BN / 2) * kGBlockK * 2));
            tma_load_2d_pair(sa, &tmap_a, full_bar(stage), kb * kGBlockK, m0);
            tma_load_2d_pair(sb, &tmap_b, full_bar(stage), kb * kGBlockK, n0);
          }
          if (++stage == Cfg::kStages) { stage = 0; phase ^= 1; }
        }
      }
      // special pieces: 128 rows of A, this CTA's half of the piece's W rows (16-row boxes), its half of the tail rows
      for (int j = 0; j < n_special; ++j) {
        if ((int)sched.owner[j] != cluster_id) continue;
        const int c0 = j * sched.sp_width;
        const int half = min(sched.sp_width, N - c0) >> 1;         // this CTA's W rows of the piece
        const int n0 = c0 + (int)rank * half;
        const uint32_t tx = 2u * (uint32_t)(Cfg::kABytes + (sched.sp_width / 2 + sched.T / 2) * kGBlockK * 2);
        for (int kb = 0; kb < k_blocks; ++kb) {
          mbar_wait(empty_bar(stage), phase ^ 1);
          const uint32_t sa = smem_base + stage * Cfg::kStageBytes;
          const uint32_t sb = sa + Cfg::kABytes;
          if (leader) mbar_arrive_expect_tx(full_bar(stage), tx);
          tma_load_2d_pair(sa, &tmap_a, full_bar(stage), kb * kGBlockK, sp_row0 + (int)rank * kGBlockM);
          // ONE box of sp_width / 2 weight rows (a narrower last piece loads rows past its own half -- of the peer's
          // half, or zero fill past N -- that no MMA column of this CTA maps to)
          tma_load_2d_pair(sb, &tmap_wsp, full_bar(stage), kb * kGBlockK, n0);
          tma_load_2d_pair(sb + tail_off, &tmap_t, full_bar(stage), kb * kGBlockK, (int)rank * (sched.T / 2));
          if (++stage == Cfg::kStages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ================================ MMA issuer (leader only) ================================
    if (leader && lane == 0) {
      constexpr uint32_t idesc = make_idesc_bf16(2 * kGBlockM, BN);
      int stage = 0;
      uint32_t phase = 0;
      int local = 0;
      for (int tile = cluster_id; tile < num_tiles; tile += num_clusters, ++local) {
        const int as = local & 1;
        const uint32_t aphase = (local >> 1) & 1;
        mbar_wait(tmem_empty_bar(as), aphase ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + as * kAccStride;
        for (int kb = 0; kb < k_blocks; ++kb) {
          mbar_wait(full_bar(stage), phase);
          tc_fence_after();
          const uint32_t sa = smem_base + stage * Cfg::kStageBytes;
          const uint32_t sb = sa + Cfg::kABytes;
          const uint64_t da = make_smem_desc(sa, 16, 1024, kLayoutSW128);
          const uint64_t db = make_smem_desc(sb, 16, 1024, kLayoutSW128);
#pragma unroll
          for (int k = 0; k < kGBlockK / 16; ++k)
            umma_f16_ss_pair(tmem_d, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), idesc, (kb > 0 || k > 0) ? 1u : 0u);
          umma_commit_pair(empty_bar(stage));
          if (++stage == Cfg::kStages) { stage = 0; phase ^= 1; }
        }
        umma_commit_pair(tmem_full_bar(as));
      }
      const uint32_t idesc_tail = make_idesc_bf16(2 * kGBlockM, sched.T);
      for (int j = 0; j < n_special; ++j) {
        if ((int)sched.owner[j] != cluster_id) continue;
        const int width = min(sched.sp_width, N - j * sched.sp_width);
        const uint32_t idesc_main = make_idesc_bf16(2 * kGBlockM, width);
        const int as = local & 1;
        const uint32_t aphase = (local >> 1) & 1;
        ++local;
        mbar_wait(tmem_empty_bar(as), aphase ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + as * kAccStride;
        for (int kb = 0; kb < k_blocks; ++kb) {
          mbar_wait(full_bar(stage), phase);
          tc_fence_after();
          const uint32_t sa = smem_base + stage * Cfg::kStageBytes;
          const uint32_t sb = sa + Cfg::kABytes;
          const uint64_t da = make_smem_desc(sa, 16, 1024, kLayoutSW128);
          const uint64_t db = make_smem_desc(sb, 16, 1024, kLayoutSW128);
          const uint64_t dt = make_smem_desc(sb + tail_off, 16, 1024, kLayoutSW128);
#pragma unroll
          for (int k = 0; k < kGBlockK / 16; ++k)
            umma_f16_ss_pair(tmem_d, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), idesc_main, (kb > 0 || k > 0) ? 1u : 0u);
          // tail rows, operands swapped: the W rows of the B slot are the M operand (rows past the piece's half are
          // whatever lies behind them: their lanes are never read back), the tail rows the N operand
#pragma unroll
          for (int k = 0; k < kGBlockK / 16; ++k)
            umma_f16_ss_pair(tmem_d + kTailCol, db + (uint64_t)(k * 2), dt + (uint64_t)(k * 2), idesc_tail, (kb > 0 || k > 0) ? 1u : 0u);
          umma_commit_pair(empty_bar(stage));
          if (++stage == Cfg::kStages) { stage = 0; phase ^= 1; }
        }
        umma_commit_pair(tmem_full_bar(as));
      }
    }
  } else {
    // ================================ epilogue (both CTAs, own 128 rows) ================================
    const int quad = warp & 3;
    const uint32_t lane_base = (uint32_t)(quad * 32) << 16;
    int local = 0;
    // rows `row` of columns [n0, n0 + width): the main accumulator of a regular tile or a special piece
    auto drain_main = [&](uint32_t taddr, int row, int n0, int width) {
      if constexpr (EPI == kEpiSwiGLU) {
        __nv_bfloat16* crow = C + (size_t)row * ldc + n0 / 2;
#pragma unroll 1
        for (int c = 0; c < width / 32; ++c) {
          uint32_t acc[32];
          tmem_ld_32x32b_x32(taddr + c * 32, acc);
          tmem_ld_wait();
          if (row < M && n0 + c * 32 < N) store_chunk_swiglu(acc, crow + c * 16);
        }
      } else {
        __nv_bfloat16* crow = C + (size_t)row * ldc + n0;
        const __nv_bfloat16* rrow = (EPI == kEpiResidual) ? R + (size_t)row * ldc + n0 : nullptr;
#pragma unroll 1
        for (int c = 0; c < width / 32; ++c) {
          uint32_t acc[32];
          tmem_ld_32x32b_x32(taddr + c * 32, acc);
          tmem_ld_wait();
          if (row < M && n0 + c * 32 < N) store_chunk<EPI>(acc, crow + c * 32, rrow + c * 32);
        }
      }
    };
    for (int tile = cluster_id; tile < num_tiles; tile += num_clusters, ++local) {
      const int as = local & 1;
      const uint32_t aphase = (local >> 1) & 1;
      const int row = (tile % m_tiles) * 2 * kGBlockM + rank * kGBlockM + quad * 32 + lane;
      mbar_wait(tmem_full_bar(as), aphase);
      tc_fence_after();
      drain_main(tmem_base + lane_base + as * kAccStride, row, (tile / m_tiles) * BN, BN);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(tmem_empty_bar(as), 0);
    }
    for (int j = 0; j < n_special; ++j) {
      if ((int)sched.owner[j] != cluster_id) continue;
      const int as = local & 1;
      const uint32_t aphase = (local >> 1) & 1;
      ++local;
      const int c0 = j * sched.sp_width;
      const int width = min(sched.sp_width, N - c0);
      mbar_wait(tmem_full_bar(as), aphase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + lane_base + as * kAccStride;
      drain_main(taddr, sp_row0 + (int)rank * kGBlockM + quad * 32 + lane, c0, width);
      // tail accumulator: TMEM lane = W row of this CTA's half (= output column), TMEM column = tail row
      const int wl = quad * 32 + lane;                              // W row inside this CTA's half
      const bool lane_ok = wl < (width >> 1);
      const int n = c0 + (int)rank * (width >> 1) + wl;             // (packed) weight row
      const int trow0 = sp_row0 + 2 * kGBlockM;
      for (int c = 0; c < sched.T / 16; ++c) {
        uint32_t acc[16];
        tmem_ld_32x32b_x16(taddr + kTailCol + c * 16, acc);
        tmem_ld_wait();
        if constexpr (EPI == kEpiSwiGLU) {
          // packed rows [gate x 16 | up x 16]: lanes 0..15 of a 32-lane group hold gate, lanes 16..31 the matching up
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const float mine = rbf(__uint_as_float(acc[i]));
            const float up = __shfl_xor_sync(0xffffffffu, mine, 16);
            const int r = c * 16 + i;
            if (lane_ok && !(lane & 16) && r < sched.tail_rows)
              C[(size_t)(trow0 + r) * ldc + (n >> 5) * 16 + (lane & 15)] = __float2bfloat16_rn(up * rbf(silu_f(mine)));
          }
        } else {
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const int r = c * 16 + i;
            if (lane_ok && r < sched.tail_rows) {
              float v = __uint_as_float(acc[i]);
              if constexpr (EPI == kEpiResidual) v = rbf(v) + __bfloat162float(R[(size_t)(trow0 + r) * ldc + n]);
              C[(size_t)(trow0 + r) * ldc + n] = __float2bfloat16_rn(v);
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(tmem_empty_bar(as), 0);
    }
  }

  tc_fence_before();
  cluster_sync_all();                     // nobody frees smem / TMEM while the pair still works
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_pair<2 * kAccStride>(tmem_base);
  }
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
int debug_gemm_flags() {   // VGPT_DEBUG_GEMM_FLAGS=1: skip the TMA loads (profiling experiments only)
  static const int v = [] { const char* e = getenv("VGPT_DEBUG_GEMM_FLAGS"); return e ? atoi(e) : 0; }();
  return v;
}

// Tile width: fewest (waves x tile width), with the narrower tile charged for its higher L2 -> SMEM fill rate per MMA
// cycle (profiles/r01c_gemm_sweep_pair.txt: qkv and gate_up prefer 256, the N = 3072 projections 192 at M = 2064).
int pick_pair_block_n(int M, int N, int num_sms) {
  const int clusters = num_sms / 2;
  const int m_tiles = (M + 255) / 256;
  double best = 0;
  int best_bn = 256;
  const int cand[2] = {256, 192};
  const double eff[2] = {1.0, 0.88};
  for (int i = 0; i < 2; ++i) {
    const int tiles = m_tiles * ((N + cand[i] - 1) / cand[i]);
    const double cost = (double)((tiles + clusters - 1) / clusters) * cand[i] / eff[i];
    if (i == 0 || cost < best - 1e-9) { best = cost; best_bn = cand[i]; }
  }
  return best_bn;
}

static int make_tmap(CUtensorMap* map, const void* base, uint64_t inner, uint64_t outer, uint64_t pitch_bytes,
                     uint32_t box_outer) {
  cuuint64_t dims[2] = {inner, outer};
  cuuint64_t strides[1] = {pitch_bytes};
  cuuint32_t box[2] = {kGBlockK, box_outer};
  cuuint32_t estr[2] = {1, 1};
  return encode_tensor_map(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                           CU_TENSOR_MAP_SWIZZLE_128B);
}

// `tail_in_loop`: M = 256 q + tail with q >= 1, 0 < tail <= kTailMax: tile rows 0 .. q-2 are regular tiles, row q-1 is cut
// into special pieces that also compute the tail rows.
template <int BN, int EPI>
static int launch_gemm(const void* A, const void* W, void* C, const void* R, int M, int N, int K, int lda, int ldc,
                       int num_sms, bool tail_in_loop, cudaStream_t stream) {
  using Cfg = GemmCfg<BN>;
  CUtensorMap ta, tb, tw, tt;
  int rc = make_tmap(&ta, A, (uint64_t)K, (uint64_t)M, (uint64_t)lda * 2, kGBlockM);
  if (rc) return rc;
  rc = make_tmap(&tb, W, (uint64_t)K, (uint64_t)N, (uint64_t)K * 2, BN / 2);
  if (rc) return rc;
  tw = tb;
  tt = ta;
  const int clusters_max = num_sms / 2;
  Sched sched;
  memset(&sched, 0, sizeof(sched));
  int m_tiles_reg = (M + 2 * kGBlockM - 1) / (2 * kGBlockM);
  const int n_tiles = (N + BN - 1) / BN;
  int work_items = m_tiles_reg * n_tiles;
  if (tail_in_loop) {
    const int tail = M % 256;
    m_tiles_reg = M / 256 - 1;
    sched.tail_rows = tail;
    // the tail MMA's N: always 32.  (N = 16 is legal but slow: with T = 16 a special piece measured ~920 cycles per
    // k-block against 448 for its 224-wide main MMAs; tools/umma_rate.py mode 4 shows N = 32 at its 39-cycle floor.)
    sched.T = kTailMax;
    // special width: main + tail accumulators in one TMEM stage, W + tail rows in the B slot; SwiGLU pairs
    // ([gate x 16 | up x 16] per 32 packed rows) must not straddle the two CTAs' halves: a multiple of 64
    int width = Cfg::kSpecialWidth;
    if (EPI == kEpiSwiGLU) width = width / 64 * 64;
    sched.sp_width = width;
    sched.n_special = (N + width - 1) / width;
    if (sched.n_special > kMaxSpecial || clusters_max > 255) {
      set_last_error("vgpt_gemm_bf16: N=%d needs %d special pieces (at most %d)", N, sched.n_special, kMaxSpecial);
      return -1;
    }
    rc = make_tmap(&tw, W, (uint64_t)K, (uint64_t)N, (uint64_t)K * 2, width / 2);
    if (rc) return rc;
    const __nv_bfloat16* a_tail = static_cast<const __nv_bfloat16*>(A) + (size_t)(M - tail) * lda;
    rc = make_tmap(&tt, a_tail, (uint64_t)K, (uint64_t)tail, (uint64_t)lda * 2, sched.T / 2);   // rows >= tail: zero fill
    if (rc) return rc;
    work_items = m_tiles_reg * n_tiles + sched.n_special;
  }
  sched.m_tiles_reg = m_tiles_reg;
  const int clusters = work_items < clusters_max ? work_items : clusters_max;
  if (tail_in_loop) {
    // List scheduling.  Regular tiles are dealt round robin; every special piece goes to the cluster that is least
    // loaded so far.  Costs in tensor-pipe cycles per k-block: a cta_group::2 MMA of width n takes n/2 cycles
    // (tools/umma_rate.py), the tail MMA about 39 (its floor).
    const int reg_tiles = m_tiles_reg * n_tiles;
    double load[256];
    for (int c = 0; c < clusters; ++c) load[c] = 2.0 * BN * ((reg_tiles + clusters - 1 - c) / clusters);
    for (int j = 0; j < sched.n_special; ++j) {
      int best = 0;
      for (int c = 1; c < clusters; ++c)
        if (load[c] < load[best] - 1e-9) best = c;
      sched.owner[j] = (uint8_t)best;
      const int w = (N - j * sched.sp_width) < sched.sp_width ? (N - j * sched.sp_width) : sched.sp_width;
      load[best] += 2.0 * w + 4.0 * 39.0;
    }
  }
  auto kern = gemm_bf16_pair_kernel<BN, EPI>;
  // per launch: the attribute is per device, and a process may drive several devices (cheap, capture-safe)
  VGPT_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes));
  kern<<<2 * clusters, kGThreads, Cfg::kSmemBytes, stream>>>(ta, tb, tw, tt, static_cast<__nv_bfloat16*>(C),
                                                            static_cast<const __nv_bfloat16*>(R), M, N, K, ldc,
                                                            debug_gemm_flags(), sched);
  VGPT_CHECK_LAUNCH();
  return 0;
}

// Modelled makespan (tensor-pipe cycles per k-block on the most loaded cluster) of a launch, plain or with the tail in
// the k-loop -- the same list scheduling launch_gemm does.  Measured (profiles/r02d_gemm_sweep_tail_in_loop.txt): the
// special pieces remove 10 % of the MMA work, but where a cluster only gets one or two work items their extra weight
// lands on the most loaded cluster (o / down at M = 2064: 2 x 384 plain, 384 + 540 special), so the choice is made per
// shape: at cfg2 qkv takes the special pieces (4 waves instead of 5), the other three projections stay plain.
static double modelled_makespan(int M, int N, int bn, int epilogue, bool in_loop, int clusters) {
  const int n_tiles = (N + bn - 1) / bn;
  if (!in_loop) {
    const int tiles = ((M + 255) / 256) * n_tiles;
    return 2.0 * bn * ((tiles + clusters - 1) / clusters);
  }
  const int reg = (M / 256 - 1) * n_tiles;
  int width = bn == 256 ? kTailCol : bn;
  if (epilogue == kEpiSwiGLU) width = width / 64 * 64;
  const int n_special = (N + width - 1) / width;
  const int c_used = reg + n_special < clusters ? reg + n_special : clusters;
  double load[256], worst = 0;
  for (int c = 0; c < c_used; ++c) load[c] = 2.0 * bn * ((reg + c_used - 1 - c) / c_used);
  for (int j = 0; j < n_special; ++j) {
    int best = 0;
    for (int c = 1; c < c_used; ++c)
      if (load[c] < load[best] - 1e-9) best = c;
    const int w = (N - j * width) < width ? (N - j * width) : width;
    load[best] += 2.0 * w + 4.0 * 39.0;
  }
  for (int c = 0; c < c_used; ++c) worst = load[c] > worst ? load[c] : worst;
  return worst;
}

// tail_mode: -1 = tuned default, 1 = plain 256-row tiles only, 3 = tail rows in the k-loop whenever the shape allows it.
int gemm_bf16(const void* A, const void* W, void* C, const void* R, int M, int N, int K, int lda, int ldc, int epilogue,
              int block_n, int tail_mode, cudaStream_t stream) {
  VGPT_CHECK_ARG(A && W && C, "vgpt_gemm_bf16: null pointer");
  VGPT_CHECK_ARG(M > 0 && N > 0 && K > 0, "vgpt_gemm_bf16: empty problem M=%d N=%d K=%d", M, N, K);
  VGPT_CHECK_ARG(K % kGBlockK == 0, "vgpt_gemm_bf16: K=%d must be a multiple of %d", K, kGBlockK);
  VGPT_CHECK_ARG(N % 64 == 0, "vgpt_gemm_bf16: N=%d must be a multiple of 64", N);
  VGPT_CHECK_ARG(lda >= K && lda % 8 == 0, "vgpt_gemm_bf16: lda=%d invalid", lda);
  VGPT_CHECK_ARG(ldc % 8 == 0, "vgpt_gemm_bf16: ldc=%d must be a multiple of 8", ldc);
  VGPT_CHECK_ARG(((uintptr_t)A & 15) == 0 && ((uintptr_t)W & 15) == 0 && ((uintptr_t)C & 15) == 0,
                 "vgpt_gemm_bf16: pointers must be 16-byte aligned");
  VGPT_CHECK_ARG(epilogue >= 0 && epilogue <= 2, "vgpt_gemm_bf16: unknown epilogue %d", epilogue);
  VGPT_CHECK_ARG(epilogue != kEpiResidual || R, "vgpt_gemm_bf16: residual epilogue needs R");
  VGPT_CHECK_ARG(tail_mode == -1 || tail_mode == 1 || tail_mode == 3, "vgpt_gemm_bf16: tail_mode %d (-1, 1 or 3)", tail_mode);
  const int sms = device_sm_count();
  // VGPT_GEMM_TAIL_IN_LOOP=0 switches the default off for A/B timing
  static const bool in_loop_default = [] { const char* e = getenv("VGPT_GEMM_TAIL_IN_LOOP"); return !(e && e[0] == '0'); }();
  const int tail = M % 256;
  const bool can = M >= 256 && tail > 0 && tail <= kTailMax && sms / 2 <= 255;
  bool in_loop = can && tail_mode == 3;
  if (block_n == 0) {
    block_n = pick_pair_block_n(M, N, sms);
    if (can && tail_mode == -1 && in_loop_default) {       // the cheaper of (plain, best width) and (special pieces, best width)
      double best = modelled_makespan(M, N, block_n, epilogue, false, sms / 2) / (block_n == 256 ? 1.0 : 0.88);
      const int cand[2] = {256, 192};
      for (int i = 0; i < 2; ++i) {
        const double t = modelled_makespan(M, N, cand[i], epilogue, true, sms / 2) / (cand[i] == 256 ? 1.0 : 0.88);
        if (t < 0.97 * best) { best = t; block_n = cand[i]; in_loop = true; }
      }
    } else if (in_loop) {
      block_n = pick_pair_block_n(M - tail, N, sms);
    }
  } else if (can && tail_mode == -1 && in_loop_default) {
    in_loop = modelled_makespan(M, N, block_n, epilogue, true, sms / 2) < 0.97 * modelled_makespan(M, N, block_n, epilogue, false, sms / 2);
  }
  VGPT_CHECK_ARG(block_n == 128 || block_n == 192 || block_n == 256, "vgpt_gemm_bf16: block_n must be 128, 192 or 256");
#define VGPT_GEMM_CASE(BN_, EPI_) \
  if (block_n == BN_ && epilogue == EPI_) return launch_gemm<BN_, EPI_>(A, W, C, R, M, N, K, lda, ldc, sms, in_loop, stream);
  VGPT_GEMM_CASE(256, kEpiStore)
  VGPT_GEMM_CASE(256, kEpiResidual)
  VGPT_GEMM_CASE(256, kEpiSwiGLU)
  VGPT_GEMM_CASE(192, kEpiStore)
  VGPT_GEMM_CASE(192, kEpiResidual)
  VGPT_GEMM_CASE(192, kEpiSwiGLU)
  VGPT_GEMM_CASE(128, kEpiStore)
  VGPT_GEMM_CASE(128, kEpiResidual)
  VGPT_GEMM_CASE(128, kEpiSwiGLU)
#undef VGPT_GEMM_CASE
  set_last_error("vgpt_gemm_bf16: no kernel for block_n=%d epilogue=%d", block_n, epilogue);
  return -1;
}

}  // namespace vgpt
