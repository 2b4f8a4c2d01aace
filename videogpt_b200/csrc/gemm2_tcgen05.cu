// bf16 GEMM on CTA pairs (tcgen05 cta_group::2): C[M,N] = A[M,K] * W[N,K]^T, same contract and
// epilogues as gemm_tcgen05.cu.
//
// Why a second kernel: with one CTA per 128 x 256 tile every SM pulls (128 + 256) x 64 bf16 per
// 512 tensor cycles = 96 B/clk from L2, and the measured ~60 % tensor-pipe activity of that kernel
// is the L2 -> SMEM fabric limit (profiles/r01a_gemm_1cta_ncu.txt).  A CTA pair computes a
// 256 x BN tile: each CTA loads its own 128 rows of A and only HALF of the W tile, and one
// tcgen05.mma.cta_group::2 (M = 256) reads both halves from the two SMs' shared memory, so the
// per-SM fill rate drops to (128 + BN/2) x 128 B per k-block (64 B/clk at BN = 256).
//
// Pair protocol (cluster of 2 along M; rank 0 = leader):
//   * both CTAs' TMA loads are .cta_group::2 and complete_tx on the LEADER's full barrier
//     (the leader's arrive.expect_tx accounts for both halves),
//   * the leader's elected thread issues the MMAs and commits with .multicast::cluster to the
//     empty / tmem-full barriers of BOTH CTAs,
//   * each CTA's epilogue warps drain their own TMEM half (128 rows) and arrive on the leader's
//     tmem-empty barrier,
//   * cluster barriers fence set-up and tear-down (the peer's smem / TMEM must outlive the
//     leader's last MMA).
#include "common.cuh"
#include "vgpt_internal.h"

#include <cuda.h>

#include <cstring>

namespace vgpt {

constexpr int kG2BlockM = 128;     // rows per CTA (256 per pair)
constexpr int kG2BlockK = 64;
constexpr int kG2Threads = 192;

template <int BN>
struct Gemm2Cfg {
  static constexpr int kABytes = kG2BlockM * kG2BlockK * 2;         // 16 KB
  static constexpr int kBBytes = (BN / 2) * kG2BlockK * 2;          // this CTA's half of the W tile
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kStages = (200 * 1024) / kStageBytes > 8 ? 8 : (200 * 1024) / kStageBytes;
  static constexpr int kTmemCols = (2 * BN <= 256) ? 256 : 512;     // two accumulator stages
  static constexpr int kBarBytes = 512;
  static constexpr int kXchgBytes = 2 * 16 * 32 * 4;               // tail tiles, SwiGLU: 16 up values x 32 lanes x 2 warp pairs
  static constexpr int kSmemBytes = kStages * kStageBytes + kBarBytes + kXchgBytes + 1024;
};

// Tail tiles (M % 256 rows that do not fill a 256-row tile), computed INSIDE the persistent kernel with the operands
// swapped:   C_tail^T[N, T] = W[N, K] * A_tail[T, K]^T
// The weight rows are the M = 256 operand of tcgen05.mma.cta_group::2 (128 per CTA, in the A slot of a stage), the T
// tail rows the N operand (T/2 per CTA, in the B slot), the accumulator holds the tail transposed (TMEM lane = output
// column, TMEM column = tail row).  A tail tile covers 256 output columns and costs about a third of a 256 x 256 main
// tile (it streams W at the L2 -> SMEM fill rate and issues N = T MMAs), instead of a whole extra row of main tiles --
// 2064 = 8 x 256 + 16 at cfg2 costs a fifth wave for qkv and an eighth for gate_up otherwise.  `owner[j]` = the cluster
// that computes tail tile j (columns [256 j, 256 j + 256)), chosen on the host by list scheduling after the main tiles.
constexpr int kMaxTailTiles = 64;
struct TailArgs {
  int rows;        // valid tail rows (0: no tail tiles)
  int T;           // rows padded to a multiple of 16 (the MMA's N), <= 128
  int n_tail;      // N / 256
  int row0;        // first tail row (= rows handled by the main tiles)
  uint8_t owner[kMaxTailTiles];
};

enum : int { kG2Store = 0, kG2Residual = 1, kG2SwiGLU = 2,
             // RMSNorm folded into the GEMMs around it (EXPERIMENTAL, vgpt_gemm_bf16_norm):
             //   (x * rstd * w) W^T == rstd * (x (W diag(w))^T)
             // 3: residual epilogue that also writes, per row and per N tile, the sum of squares of the
             //    bf16 values it stores (fixed slots, no atomics: the consumer's sum is deterministic);
             // 4 / 5: store / SwiGLU epilogues that scale every accumulator row by
             //    rstd = rsqrt(sum of that row's parts / K + eps) before the bf16 rounding.
             kG2ResidualSS = 3, kG2StoreScaled = 4, kG2SwiGLUScaled = 5 };

constexpr int kNormParts = 32;     // slots per row of the sum-of-squares buffer [M][kNormParts]
struct NormArgs { float* ss; float inv_k; float eps; };
template <int EPI> struct EpiExtra { };                                  // nothing for the plain epilogues
template <> struct EpiExtra<kG2ResidualSS> { NormArgs na; };
template <> struct EpiExtra<kG2StoreScaled> { NormArgs na; };
template <> struct EpiExtra<kG2SwiGLUScaled> { NormArgs na; };

__device__ __forceinline__ void tmem_ld_32x32b_x16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
// One 32-column chunk of one accumulator row -> global memory (same as the 1-CTA kernel).
template <int EPI>
__device__ __forceinline__ void store_chunk2(const uint32_t (&acc)[32], __nv_bfloat16* __restrict__ out,
                                             const __nv_bfloat16* __restrict__ res) {
  uint4 r[4];
  if constexpr (EPI == kG2Residual) {
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
      if constexpr (EPI == kG2Residual) {
        uint32_t rv = (&r[i].x)[j];
        a = rbf(a) + bf16lo(rv);
        b = rbf(b) + bf16hi(rv);
      }
      w[j] = pack_bf16x2(a, b);
    }
    op[i] = make_uint4(w[0], w[1], w[2], w[3]);
  }
}

// Residual chunk that also returns the sum of squares of the 32 bf16 values it stores.
__device__ __forceinline__ float store_chunk2_ss(const uint32_t (&acc)[32], __nv_bfloat16* __restrict__ out,
                                                 const __nv_bfloat16* __restrict__ res) {
  uint4 r[4];
  const uint4* rp = reinterpret_cast<const uint4*>(res);
#pragma unroll
  for (int i = 0; i < 4; ++i) r[i] = rp[i];
  uint4* op = reinterpret_cast<uint4*>(out);
  float ss = 0.f;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    uint32_t w[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const uint32_t rv = (&r[i].x)[j];
      const float a = rbf(rbf(__uint_as_float(acc[i * 8 + j * 2])) + bf16lo(rv));
      const float b = rbf(rbf(__uint_as_float(acc[i * 8 + j * 2 + 1])) + bf16hi(rv));
      ss = fmaf(a, a, ss);
      ss = fmaf(b, b, ss);
      w[j] = pack_bf16x2(a, b);
    }
    op[i] = make_uint4(w[0], w[1], w[2], w[3]);
  }
  return ss;
}

template <int BN, int EPI>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kG2Threads, 1)
gemm_bf16_tcgen05_pair_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                              const __grid_constant__ CUtensorMap tmap_w, const __grid_constant__ CUtensorMap tmap_t,
                              __nv_bfloat16* __restrict__ C, const __nv_bfloat16* __restrict__ R, int M, int N,
                              int K, int ldc, int flags, EpiExtra<EPI> ex, const __grid_constant__ TailArgs tail) {
  using Cfg = Gemm2Cfg<BN>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bar_base = smem_base + Cfg::kStages * Cfg::kStageBytes;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (Cfg::kStages + s); };
  auto tmem_full_bar = [&](int s) { return bar_base + 8u * (2 * Cfg::kStages + s); };
  auto tmem_empty_bar = [&](int s) { return bar_base + 8u * (2 * Cfg::kStages + 2 + s); };
  const uint32_t tmem_slot = bar_base + 8u * (2 * Cfg::kStages + 4);
  uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));
  [[maybe_unused]] float* xchg = reinterpret_cast<float*>(smem_raw + (bar_base + Cfg::kBarBytes - smem_u32(smem_raw)));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int cluster_id = blockIdx.x >> 1, num_clusters = gridDim.x >> 1;
  const int m_tiles = (M + 2 * kG2BlockM - 1) / (2 * kG2BlockM);
  const int n_tiles = (N + BN - 1) / BN;
  const int num_tiles = m_tiles * n_tiles;
  const int k_blocks = K / kG2BlockK;
  const int n_tail = tail.rows > 0 ? tail.n_tail : 0;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_b);
    if (n_tail) { tma_prefetch_desc(&tmap_w); tma_prefetch_desc(&tmap_t); }
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
  if (warp == 1) tmem_alloc_pair<Cfg::kTmemCols>(tmem_slot);
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
        const int m0 = (tile % m_tiles) * 2 * kG2BlockM + rank * kG2BlockM;
        const int n0 = (tile / m_tiles) * BN + rank * (BN / 2);
        for (int kb = 0; kb < k_blocks; ++kb) {
          mbar_wait(empty_bar(stage), phase ^ 1);
          const uint32_t sa = smem_base + stage * Cfg::kStageBytes;
          const uint32_t sb = sa + Cfg::kABytes;
          if (flags & 1) {                 // debug: MMA-bound ceiling, operands not refreshed
            if (leader) mbar_arrive(full_bar(stage));
          } else {
            // The peer only issues its loads: its bytes are credited to the leader's barrier, whose
            // phase cannot complete before the leader's own arrive.expect_tx (count 1).  (A remote
            // mbarrier.arrive.release.cluster here serialised the peer's loads: profiles/r01c.)
            if (leader) mbar_arrive_expect_tx(full_bar(stage), 2 * Cfg::kStageBytes);
            tma_load_2d_pair(sa, &tmap_a, full_bar(stage), kb * kG2BlockK, m0);
            tma_load_2d_pair(sb, &tmap_b, full_bar(stage), kb * kG2BlockK, n0);
          }
          if (++stage == Cfg::kStages) { stage = 0; phase ^= 1; }
        }
      }
      // tail tiles: this CTA's 128 weight rows into the A slot, its half of the tail rows into the B slot
      const uint32_t tail_tx = 2u * (uint32_t)(Cfg::kABytes + (tail.T / 2) * kG2BlockK * 2);
      for (int j = 0; j < n_tail; ++j) {
        if ((int)tail.owner[j] != cluster_id) continue;
        const int n0 = j * 2 * kG2BlockM + rank * kG2BlockM;
        for (int kb = 0; kb < k_blocks; ++kb) {
          mbar_wait(empty_bar(stage), phase ^ 1);
          const uint32_t sa = smem_base + stage * Cfg::kStageBytes;
          const uint32_t sb = sa + Cfg::kABytes;
          if (leader) mbar_arrive_expect_tx(full_bar(stage), tail_tx);
          tma_load_2d_pair(sa, &tmap_w, full_bar(stage), kb * kG2BlockK, n0);
          tma_load_2d_pair(sb, &tmap_t, full_bar(stage), kb * kG2BlockK, (int)rank * (tail.T / 2));
          if (++stage == Cfg::kStages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ================================ MMA issuer (leader only) ================================
    if (leader && lane == 0) {
      constexpr uint32_t idesc = make_idesc_bf16(2 * kG2BlockM, BN);
      int stage = 0;
      uint32_t phase = 0;
      int local = 0;
      for (int tile = cluster_id; tile < num_tiles; tile += num_clusters, ++local) {
        const int as = local & 1;
        const uint32_t aphase = (local >> 1) & 1;
        mbar_wait(tmem_empty_bar(as), aphase ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + as * BN;
        for (int kb = 0; kb < k_blocks; ++kb) {
          mbar_wait(full_bar(stage), phase);
          tc_fence_after();
          const uint32_t sa = smem_base + stage * Cfg::kStageBytes;
          const uint32_t sb = sa + Cfg::kABytes;
          const uint64_t da = make_smem_desc(sa, 16, 1024, kLayoutSW128);
          const uint64_t db = make_smem_desc(sb, 16, 1024, kLayoutSW128);
#pragma unroll
          for (int k = 0; k < kG2BlockK / 16; ++k)
            umma_f16_ss_pair(tmem_d, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), idesc, (kb > 0 || k > 0) ? 1u : 0u);
          umma_commit_pair(empty_bar(stage));
          if (++stage == Cfg::kStages) { stage = 0; phase ^= 1; }
        }
        umma_commit_pair(tmem_full_bar(as));
      }
      const uint32_t idesc_tail = make_idesc_bf16(2 * kG2BlockM, tail.T);
      for (int j = 0; j < n_tail; ++j) {
        if ((int)tail.owner[j] != cluster_id) continue;
        const int as = local & 1;
        const uint32_t aphase = (local >> 1) & 1;
        ++local;
        mbar_wait(tmem_empty_bar(as), aphase ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + as * BN;
        for (int kb = 0; kb < k_blocks; ++kb) {
          mbar_wait(full_bar(stage), phase);
          tc_fence_after();
          const uint32_t sa = smem_base + stage * Cfg::kStageBytes;
          const uint64_t da = make_smem_desc(sa, 16, 1024, kLayoutSW128);
          const uint64_t db = make_smem_desc(sa + Cfg::kABytes, 16, 1024, kLayoutSW128);
#pragma unroll
          for (int k = 0; k < kG2BlockK / 16; ++k)
            umma_f16_ss_pair(tmem_d, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), idesc_tail, (kb > 0 || k > 0) ? 1u : 0u);
          umma_commit_pair(empty_bar(stage));
          if (++stage == Cfg::kStages) { stage = 0; phase ^= 1; }
        }
        umma_commit_pair(tmem_full_bar(as));
      }
    }
  } else {
    // ================================ epilogue (both CTAs, own 128 rows) ================================
    const int quad = warp & 3;
    int local = 0;
    for (int tile = cluster_id; tile < num_tiles; tile += num_clusters, ++local) {
      const int as = local & 1;
      const uint32_t aphase = (local >> 1) & 1;
      const int m0 = (tile % m_tiles) * 2 * kG2BlockM + rank * kG2BlockM;
      const int n0 = (tile / m_tiles) * BN;
      const int row = m0 + quad * 32 + lane;
      mbar_wait(tmem_full_bar(as), aphase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + as * BN;
      [[maybe_unused]] float rstd = 1.f;
      if constexpr (EPI == kG2StoreScaled || EPI == kG2SwiGLUScaled) {
        if (row < M) {                       // fixed-order sum of the producer's per-tile parts
          const float4* p = reinterpret_cast<const float4*>(ex.na.ss + (size_t)row * kNormParts);
          float ssum = 0.f;
#pragma unroll
          for (int i = 0; i < kNormParts / 4; ++i) { const float4 v = p[i]; ssum += v.x; ssum += v.y; ssum += v.z; ssum += v.w; }
          rstd = rsqrtf(ssum * ex.na.inv_k + ex.na.eps);
        }
      }
      if constexpr (EPI == kG2SwiGLU || EPI == kG2SwiGLUScaled) {
        __nv_bfloat16* crow = C + (size_t)row * ldc + n0 / 2;
#pragma unroll 1
        for (int c = 0; c < BN / 64; ++c) {
          uint32_t g[32], u[32];
          tmem_ld_32x32b_x32(taddr + c * 64, g);
          tmem_ld_32x32b_x32(taddr + c * 64 + 32, u);
          tmem_ld_wait();
          if (row < M && n0 + c * 64 < N) {
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              float gv, uv;
              if constexpr (EPI == kG2SwiGLUScaled) {
                gv = rbf(__uint_as_float(g[i]) * rstd);
                uv = rbf(__uint_as_float(u[i]) * rstd);
              } else {
                gv = rbf(__uint_as_float(g[i]));
                uv = rbf(__uint_as_float(u[i]));
              }
              g[i] = __float_as_uint(uv * rbf(silu_f(gv)));
            }
            store_chunk2<kG2Store>(g, crow + c * 32, nullptr);
          }
        }
      } else if constexpr (EPI == kG2ResidualSS) {
        __nv_bfloat16* crow = C + (size_t)row * ldc + n0;
        const __nv_bfloat16* rrow = R + (size_t)row * ldc + n0;
        float ss = 0.f;
#pragma unroll 1
        for (int c = 0; c < BN / 32; ++c) {
          uint32_t acc[32];
          tmem_ld_32x32b_x32(taddr + c * 32, acc);
          tmem_ld_wait();
          if (row < M && n0 + c * 32 < N) ss += store_chunk2_ss(acc, crow + c * 32, rrow + c * 32);
        }
        if (row < M) {
          const int part = tile / m_tiles;           // this tile's slot; slot 0's owner clears the unused ones
          float* p = ex.na.ss + (size_t)row * kNormParts;
          p[part] = ss;
          if (part == 0)
            for (int i = n_tiles; i < kNormParts; ++i) p[i] = 0.f;
        }
      } else {
        __nv_bfloat16* crow = C + (size_t)row * ldc + n0;
        const __nv_bfloat16* rrow = (EPI == kG2Residual) ? R + (size_t)row * ldc + n0 : nullptr;
#pragma unroll 1
        for (int c = 0; c < BN / 32; ++c) {
          uint32_t acc[32];
          tmem_ld_32x32b_x32(taddr + c * 32, acc);
          tmem_ld_wait();
          if constexpr (EPI == kG2StoreScaled) {
#pragma unroll
            for (int i = 0; i < 32; ++i) acc[i] = __float_as_uint(__uint_as_float(acc[i]) * rstd);
            if (row < M && n0 + c * 32 < N) store_chunk2<kG2Store>(acc, crow + c * 32, nullptr);
          } else {
            if (row < M && n0 + c * 32 < N) store_chunk2<EPI>(acc, crow + c * 32, rrow + c * 32);
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(tmem_empty_bar(as), 0);
    }
    // ---- tail tiles: TMEM lane = output column (weight row), TMEM column = tail row --------------------
    if constexpr (EPI == kG2Store || EPI == kG2Residual || EPI == kG2SwiGLU) {
      for (int j = 0; j < n_tail; ++j) {
        if ((int)tail.owner[j] != cluster_id) continue;
        const int as = local & 1;
        const uint32_t aphase = (local >> 1) & 1;
        ++local;
        mbar_wait(tmem_full_bar(as), aphase);
        tc_fence_after();
        const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + as * BN;
        const int n = j * 2 * kG2BlockM + (int)rank * kG2BlockM + quad * 32 + lane;     // weight row
        for (int c = 0; c < tail.T / 16; ++c) {
          uint32_t acc[16];
          tmem_ld_32x32b_x16(taddr + c * 16, acc);
          tmem_ld_wait();
          if constexpr (EPI == kG2SwiGLU) {
            // packed weight rows: [gate x 32 | up x 32] per 64 -> even quads hold gate, odd quads the matching up
            float* x = xchg + (quad >> 1) * 16 * 32;
            if (quad & 1) {
#pragma unroll
              for (int i = 0; i < 16; ++i) x[i * 32 + lane] = __uint_as_float(acc[i]);
            }
            asm volatile("bar.sync %0, 64;" ::"r"(1 + (quad >> 1)) : "memory");
            if (!(quad & 1)) {
              const int oc = (j * 2 * kG2BlockM + (int)rank * kG2BlockM) / 2 + (quad >> 1) * 32 + lane;
#pragma unroll
              for (int i = 0; i < 16; ++i) {
                const int r = c * 16 + i;
                if (r < tail.rows) {
                  const float gv = rbf(__uint_as_float(acc[i]));
                  const float uv = rbf(x[i * 32 + lane]);
                  C[(size_t)(tail.row0 + r) * ldc + oc] = __float2bfloat16_rn(uv * rbf(silu_f(gv)));
                }
              }
            }
            asm volatile("bar.sync %0, 64;" ::"r"(1 + (quad >> 1)) : "memory");      // x reusable
          } else {
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              const int r = c * 16 + i;
              if (r < tail.rows) {
                float v = __uint_as_float(acc[i]);
                if constexpr (EPI == kG2Residual) v = rbf(v) + __bfloat162float(R[(size_t)(tail.row0 + r) * ldc + n]);
                C[(size_t)(tail.row0 + r) * ldc + n] = __float2bfloat16_rn(v);
              }
            }
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(tmem_empty_bar(as), 0);
      }
    }
  }

  tc_fence_before();
  cluster_sync_all();                     // nobody frees smem / TMEM while the pair still works
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_pair<Cfg::kTmemCols>(tmem_base);
  }
}

// ---------------------------------------------------------------------------------------------
// Skinny tail GEMM: the last M % 256 <= T rows of a projection (the 16 tag / time-slot rows that
// make M = 2064 = 8 x 256 + 16 at cfg2 would otherwise cost a ninth, almost empty row of 256-row
// tiles -- 11 % of all tiles and, for qkv / gate_up, a whole extra wave).  Operands swapped:
//     C_tail^T[N, T] = W[N, K] * A_tail[T, K]^T
// so the WEIGHT tile is the M = 256 operand of tcgen05.mma.cta_group::2 (128 rows per CTA), the T
// tail rows are the N = T operand (T/2 rows per CTA), and the accumulator holds the tail
// transposed (TMEM lane = output column, TMEM column = tail row).  The MMAs are tiny (N = 16);
// the kernel streams W once at the L2 -> SMEM fill rate.  Same epilogues; for SwiGLU the gate and
// up halves of an output land in neighbouring warps and are exchanged through shared memory.
// ---------------------------------------------------------------------------------------------
template <int T>
struct SkinnyCfg {
  static constexpr int kABytes = kG2BlockM * kG2BlockK * 2;          // 128 weight rows x 64 k
  static constexpr int kBBytes = (T / 2) * kG2BlockK * 2;            // this CTA's half of the tail rows
  static constexpr int kStageBytes = kABytes + ((kBBytes + 1023) / 1024) * 1024;
  static constexpr int kStages = 8;
  static constexpr int kTmemCols = (2 * T < 32) ? 32 : 2 * T;
  static constexpr int kXchgBytes = 2 * T * 32 * 4;                  // SwiGLU: up values of two warp pairs
  static constexpr int kSmemBytes = kStages * kStageBytes + 512 + kXchgBytes + 1024;
};

template <int T, int EPI>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kG2Threads, 1)
gemm_bf16_skinny_pair_kernel(const __grid_constant__ CUtensorMap tmap_w, const __grid_constant__ CUtensorMap tmap_a,
                             __nv_bfloat16* __restrict__ C, const __nv_bfloat16* __restrict__ R, int rows, int N,
                             int K, int ldc) {
  using Cfg = SkinnyCfg<T>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bar_base = smem_base + Cfg::kStages * Cfg::kStageBytes;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (Cfg::kStages + s); };
  auto tmem_full_bar = [&](int s) { return bar_base + 8u * (2 * Cfg::kStages + s); };
  auto tmem_empty_bar = [&](int s) { return bar_base + 8u * (2 * Cfg::kStages + 2 + s); };
  const uint32_t tmem_slot = bar_base + 8u * (2 * Cfg::kStages + 4);
  uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));
  float* xchg = reinterpret_cast<float*>(smem_raw + (bar_base + 512 - smem_u32(smem_raw)));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int cluster_id = blockIdx.x >> 1, num_clusters = gridDim.x >> 1;
  const int num_tiles = N / (2 * kG2BlockM);          // 256 output columns per tile
  const int k_blocks = K / kG2BlockK;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_w);
    tma_prefetch_desc(&tmap_a);
    for (int s = 0; s < Cfg::kStages; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(tmem_full_bar(s), 1); mbar_init(tmem_empty_bar(s), 8); }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc_pair<Cfg::kTmemCols>(tmem_slot);
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  if (warp == 0) {
    // ================================ TMA producer (both CTAs) ================================
    int stage = 0;
    uint32_t phase = 0;
    for (int tile = cluster_id; tile < num_tiles; tile += num_clusters) {
      const int n0 = tile * 2 * kG2BlockM + rank * kG2BlockM;         // this CTA's 128 weight rows
      for (int kb = 0; kb < k_blocks; ++kb) {
        mbar_wait(empty_bar(stage), phase ^ 1);
        if (elect_one_sync()) {
          const uint32_t sa = smem_base + stage * Cfg::kStageBytes;
          const uint32_t sb = sa + Cfg::kABytes;
          if (leader) mbar_arrive_expect_tx(full_bar(stage), 2 * (Cfg::kABytes + Cfg::kBBytes));
          tma_load_2d_pair(sa, &tmap_w, full_bar(stage), kb * kG2BlockK, n0);
          tma_load_2d_pair(sb, &tmap_a, full_bar(stage), kb * kG2BlockK, (int)rank * (T / 2));
        }
        __syncwarp();
        if (++stage == Cfg::kStages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ================================ MMA issuer (leader only) ================================
    if (leader) {
      constexpr uint32_t idesc = make_idesc_bf16(2 * kG2BlockM, T);
      int stage = 0;
      uint32_t phase = 0;
      int local = 0;
      for (int tile = cluster_id; tile < num_tiles; tile += num_clusters, ++local) {
        const int as = local & 1;
        mbar_wait(tmem_empty_bar(as), ((local >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + as * T;
        for (int kb = 0; kb < k_blocks; ++kb) {
          mbar_wait(full_bar(stage), phase);
          tc_fence_after();
          if (elect_one_sync()) {
            const uint32_t sa = smem_base + stage * Cfg::kStageBytes;
            const uint64_t da = make_smem_desc(sa, 16, 1024, kLayoutSW128);
            const uint64_t db = make_smem_desc(sa + Cfg::kABytes, 16, 1024, kLayoutSW128);
#pragma unroll
            for (int k = 0; k < kG2BlockK / 16; ++k)
              umma_f16_ss_pair(tmem_d, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), idesc, (kb > 0 || k > 0) ? 1u : 0u);
            umma_commit_pair(empty_bar(stage));
            if (kb == k_blocks - 1) umma_commit_pair(tmem_full_bar(as));
          }
          __syncwarp();
          if (++stage == Cfg::kStages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else {
    // ============== epilogue (both CTAs): TMEM lane = output column, column = tail row ==============
    const int quad = warp & 3;
    int local = 0;
    for (int tile = cluster_id; tile < num_tiles; tile += num_clusters, ++local) {
      const int as = local & 1;
      mbar_wait(tmem_full_bar(as), (local >> 1) & 1);
      tc_fence_after();
      uint32_t acc[T];
      const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + as * T;
      if constexpr (T == 16) {
        tmem_ld_32x32b_x16(taddr, acc);
      } else {
        tmem_ld_32x32b_x32(taddr, acc);
      }
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(tmem_empty_bar(as), 0);      // accumulator drained
      const int n = tile * 2 * kG2BlockM + (int)rank * kG2BlockM + quad * 32 + lane;   // weight row = output column
      if constexpr (EPI == kG2SwiGLU) {
        // packed rows: [gate x 32 | up x 32] per 64 -> even warps hold gate, odd warps the matching up
        float* x = xchg + (quad >> 1) * T * 32;
        if (quad & 1) {
#pragma unroll
          for (int j = 0; j < T; ++j) x[j * 32 + lane] = __uint_as_float(acc[j]);
        }
        asm volatile("bar.sync %0, 64;" ::"r"(1 + (quad >> 1)) : "memory");
        if (!(quad & 1)) {
          const int oc = (tile * 2 * kG2BlockM + (int)rank * kG2BlockM) / 2 + (quad >> 1) * 32 + lane;
#pragma unroll
          for (int j = 0; j < T; ++j) {
            if (j < rows) {
              const float gv = rbf(__uint_as_float(acc[j]));
              const float uv = rbf(x[j * 32 + lane]);
              C[(size_t)j * ldc + oc] = __float2bfloat16_rn(uv * rbf(silu_f(gv)));
            }
          }
        }
        asm volatile("bar.sync %0, 64;" ::"r"(1 + (quad >> 1)) : "memory");      // x reusable for the next tile
      } else {
#pragma unroll
        for (int j = 0; j < T; ++j) {
          if (j < rows) {
            float v = __uint_as_float(acc[j]);
            if constexpr (EPI == kG2Residual) v = rbf(v) + __bfloat162float(R[(size_t)j * ldc + n]);
            C[(size_t)j * ldc + n] = __float2bfloat16_rn(v);
          }
        }
      }
    }
  }

  tc_fence_before();
  cluster_sync_all();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_pair<Cfg::kTmemCols>(tmem_base);
  }
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
// `tail_rows` > 0: rows [M, M + tail_rows) of A / C / R are computed by swapped-operand tail tiles inside the same
// launch (M is then a multiple of 256, or 0); needs N % 256 == 0 and a plain epilogue (0..2).
template <int BN, int EPI>
static int launch_gemm2(const void* A, const void* W, void* C, const void* R, int M, int N, int K, int lda,
                        int ldc, int num_sms, cudaStream_t stream, EpiExtra<EPI> ex = EpiExtra<EPI>(), int tail_rows = 0) {
  using Cfg = Gemm2Cfg<BN>;
  CUtensorMap ta, tb, tw, tt;
  cuuint32_t estr[2] = {1, 1};
  {
    cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)(M > 0 ? M : 1)};
    cuuint64_t strides[1] = {(cuuint64_t)lda * 2};
    cuuint32_t box[2] = {kG2BlockK, kG2BlockM};
    int rc = encode_tensor_map(&ta, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(A), dims, strides, box,
                               estr, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
  }
  {
    cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)N};
    cuuint64_t strides[1] = {(cuuint64_t)K * 2};
    cuuint32_t box[2] = {kG2BlockK, BN / 2};
    int rc = encode_tensor_map(&tb, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(W), dims, strides, box,
                               estr, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
  }
  const int clusters_max = num_sms / 2;
  const int tiles = ((M + 2 * kG2BlockM - 1) / (2 * kG2BlockM)) * ((N + BN - 1) / BN);
  TailArgs tail;
  memset(&tail, 0, sizeof(tail));
  tw = tb;
  tt = ta;
  int clusters = tiles < clusters_max ? tiles : clusters_max;
  if (tail_rows > 0) {
    tail.rows = tail_rows;
    tail.T = (tail_rows + 15) & ~15;
    tail.n_tail = N / (2 * kG2BlockM);
    tail.row0 = M;
    {
      cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)N};
      cuuint64_t strides[1] = {(cuuint64_t)K * 2};
      cuuint32_t box[2] = {kG2BlockK, kG2BlockM};
      int rc = encode_tensor_map(&tw, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(W), dims, strides, box,
                                 estr, CU_TENSOR_MAP_SWIZZLE_128B);
      if (rc) return rc;
    }
    {
      cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)tail_rows};          // rows beyond tail_rows are zero-filled
      cuuint64_t strides[1] = {(cuuint64_t)lda * 2};
      cuuint32_t box[2] = {kG2BlockK, (cuuint32_t)(tail.T / 2)};
      const __nv_bfloat16* a_tail = static_cast<const __nv_bfloat16*>(A) + (size_t)M * lda;
      int rc = encode_tensor_map(&tt, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<__nv_bfloat16*>(a_tail), dims, strides,
                                 box, estr, CU_TENSOR_MAP_SWIZZLE_128B);
      if (rc) return rc;
    }
    // list scheduling: main tiles are dealt round robin (tile t -> cluster t % clusters); every tail tile goes to the
    // cluster that is least loaded so far.  Costs in tensor-pipe cycles per k-block: a main tile issues 4 MMAs of
    // N/2 + 43 cycles (SS form, profiles/r01h_umma_rate.txt); a tail tile 4 MMAs of T/2 + 43, but not less than the
    // ~200 cycles its 17 KB per SM take through the L2 -> SMEM fabric.
    clusters = (tiles + tail.n_tail) < clusters_max ? (tiles + tail.n_tail) : clusters_max;
    const double c_main = 4.0 * (BN / 2 + 43);
    double c_tail = 4.0 * (tail.T / 2 + 43);
    if (c_tail < 200.0) c_tail = 200.0;
    double load[256];
    for (int c = 0; c < clusters; ++c) load[c] = c_main * ((tiles + clusters - 1 - c) / clusters);
    for (int j = 0; j < tail.n_tail; ++j) {
      int best = 0;
      for (int c = 1; c < clusters; ++c)
        if (load[c] < load[best] - 1e-9) best = c;
      tail.owner[j] = (uint8_t)best;
      load[best] += c_tail;
    }
  }
  auto kern = gemm_bf16_tcgen05_pair_kernel<BN, EPI>;
  // per launch: the attribute is per device, and a process may drive several devices (cheap, capture-safe)
  VGPT_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes));
  kern<<<2 * clusters, kG2Threads, Cfg::kSmemBytes, stream>>>(ta, tb, tw, tt, static_cast<__nv_bfloat16*>(C),
                                                             static_cast<const __nv_bfloat16*>(R), M, N, K, ldc,
                                                             debug_gemm_flags(), ex, tail);
  VGPT_CHECK_LAUNCH();
  return 0;
}


template <int T, int EPI>
static int launch_skinny(const void* A_tail, const void* W, void* C_tail, const void* R_tail, int rows, int N, int K,
                         int lda, int ldc, int num_sms, cudaStream_t stream) {
  using Cfg = SkinnyCfg<T>;
  CUtensorMap tw, ta;
  {
    cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)N};
    cuuint64_t strides[1] = {(cuuint64_t)K * 2};
    cuuint32_t box[2] = {kG2BlockK, kG2BlockM}, estr[2] = {1, 1};
    int rc = encode_tensor_map(&tw, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(W), dims, strides, box,
                               estr, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
  }
  {
    cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)rows};           // rows beyond `rows` are zero-filled
    cuuint64_t strides[1] = {(cuuint64_t)lda * 2};
    cuuint32_t box[2] = {kG2BlockK, T / 2}, estr[2] = {1, 1};
    int rc = encode_tensor_map(&ta, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(A_tail), dims, strides, box,
                               estr, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
  }
  auto kern = gemm_bf16_skinny_pair_kernel<T, EPI>;
  VGPT_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes));
  const int tiles = N / (2 * kG2BlockM);
  const int clusters = tiles < num_sms / 2 ? tiles : num_sms / 2;
  kern<<<2 * clusters, kG2Threads, Cfg::kSmemBytes, stream>>>(tw, ta, static_cast<__nv_bfloat16*>(C_tail),
                                                             static_cast<const __nv_bfloat16*>(R_tail), rows, N, K, ldc);
  VGPT_CHECK_LAUNCH();
  return 0;
}

// rows <= 32 tail rows of a projection (A_tail / C_tail / R_tail point at the first tail row).
int gemm_bf16_skinny(const void* A_tail, const void* W, void* C_tail, const void* R_tail, int rows, int N, int K,
                     int lda, int ldc, int epilogue, cudaStream_t stream) {
  VGPT_CHECK_ARG(rows > 0 && rows <= 32 && N % 256 == 0 && K % kG2BlockK == 0,
                 "vgpt_gemm_bf16: skinny tail needs rows <= 32, N %% 256 == 0 (rows=%d N=%d)", rows, N);
  const int sms = device_sm_count();
#define VGPT_SKINNY_CASE(T_, EPI_) \
  if ((rows <= 16) == (T_ == 16) && epilogue == EPI_) \
    return launch_skinny<T_, EPI_>(A_tail, W, C_tail, R_tail, rows, N, K, lda, ldc, sms, stream);
  VGPT_SKINNY_CASE(16, kG2Store)
  VGPT_SKINNY_CASE(16, kG2Residual)
  VGPT_SKINNY_CASE(16, kG2SwiGLU)
  VGPT_SKINNY_CASE(32, kG2Store)
  VGPT_SKINNY_CASE(32, kG2Residual)
  VGPT_SKINNY_CASE(32, kG2SwiGLU)
#undef VGPT_SKINNY_CASE
  set_last_error("vgpt_gemm_bf16: no skinny kernel for epilogue=%d", epilogue);
  return -1;
}

// RMSNorm folded into the neighbouring GEMMs (EXPERIMENTAL: never run on hardware).  epilogue 3 writes the
// per-row, per-N-tile sums of squares into row_ss[M][kNormParts]; 4 / 5 read them.  CTA pairs only.
int gemm_bf16_norm(const void* A, const void* W, void* C, const void* R, int M, int N, int K, int lda, int ldc,
                   int epilogue, float* row_ss, float eps, cudaStream_t stream) {
  VGPT_CHECK_ARG(A && W && C && row_ss && M > 0 && N > 0 && K > 0 && K % kG2BlockK == 0 && N % 64 == 0,
                 "vgpt_gemm_bf16_norm: bad arguments (M=%d N=%d K=%d)", M, N, K);
  VGPT_CHECK_ARG(epilogue >= kG2ResidualSS && epilogue <= kG2SwiGLUScaled, "vgpt_gemm_bf16_norm: epilogue %d", epilogue);
  VGPT_CHECK_ARG(epilogue != kG2ResidualSS || R, "vgpt_gemm_bf16_norm: the residual epilogue needs R");
  VGPT_CHECK_ARG(((uintptr_t)row_ss & 15) == 0, "vgpt_gemm_bf16_norm: row_ss must be 16-byte aligned");
  const int sms = device_sm_count();
  const int bn = pick_pair_block_n(M, N, sms);
  VGPT_CHECK_ARG(epilogue != kG2ResidualSS || (N + bn - 1) / bn <= kNormParts,
                 "vgpt_gemm_bf16_norm: N=%d needs more than %d sum-of-squares slots", N, kNormParts);
  NormArgs na{row_ss, 1.0f / (float)K, eps};
#define VGPT_GEMM2N_CASE(BN_, EPI_)                                                                   \
  if (bn == BN_ && epilogue == EPI_) {                                                                \
    EpiExtra<EPI_> ex;                                                                                \
    ex.na = na;                                                                                       \
    return launch_gemm2<BN_, EPI_>(A, W, C, R, M, N, K, lda, ldc, sms, stream, ex);                  \
  }
  VGPT_GEMM2N_CASE(256, kG2ResidualSS)
  VGPT_GEMM2N_CASE(192, kG2ResidualSS)
  VGPT_GEMM2N_CASE(256, kG2StoreScaled)
  VGPT_GEMM2N_CASE(192, kG2StoreScaled)
  VGPT_GEMM2N_CASE(256, kG2SwiGLUScaled)
  VGPT_GEMM2N_CASE(192, kG2SwiGLUScaled)
#undef VGPT_GEMM2N_CASE
  set_last_error("vgpt_gemm_bf16_norm: no kernel for block_n=%d epilogue=%d", bn, epilogue);
  return -1;
}

// `tail_rows` > 0: M main rows (a multiple of 256, or 0) + tail_rows <= 128 more rows computed by tail tiles.
int gemm_bf16_pair(const void* A, const void* W, void* C, const void* R, int M, int N, int K, int lda, int ldc,
                   int epilogue, int block_n, cudaStream_t stream, int tail_rows) {
  const int sms = device_sm_count();
  VGPT_CHECK_ARG(tail_rows >= 0 && tail_rows <= 128 && (tail_rows == 0 || (M % 256 == 0 && N % 256 == 0 && N / 256 <= kMaxTailTiles && sms / 2 <= 255)),
                 "vgpt_gemm_bf16: tail tiles need M %% 256 == 0, N %% 256 == 0, N <= %d (M=%d N=%d tail=%d)", 256 * kMaxTailTiles, M, N, tail_rows);
#define VGPT_GEMM2_CASE(BN_, EPI_) \
  if (block_n == BN_ && epilogue == EPI_) \
    return launch_gemm2<BN_, EPI_>(A, W, C, R, M, N, K, lda, ldc, sms, stream, EpiExtra<EPI_>(), tail_rows);
  VGPT_GEMM2_CASE(256, kG2Store)
  VGPT_GEMM2_CASE(256, kG2Residual)
  VGPT_GEMM2_CASE(256, kG2SwiGLU)
  VGPT_GEMM2_CASE(192, kG2Store)
  VGPT_GEMM2_CASE(192, kG2Residual)
  VGPT_GEMM2_CASE(192, kG2SwiGLU)
  VGPT_GEMM2_CASE(128, kG2Store)
  VGPT_GEMM2_CASE(128, kG2Residual)
  VGPT_GEMM2_CASE(128, kG2SwiGLU)
#undef VGPT_GEMM2_CASE
  set_last_error("vgpt_gemm_bf16: no CTA-pair kernel for block_n=%d epilogue=%d", block_n, epilogue);
  return -1;
}

}  // namespace vgpt
