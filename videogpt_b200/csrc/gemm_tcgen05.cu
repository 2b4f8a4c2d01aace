// bf16 GEMM  C[M,N] = A[M,K] * W[N,K]^T  for the QKV / O / gate-up / down projections of the
// Phi-3 block (reference call sites: LVM/transform/sdpa_transform.py:39,89 and transformers
// Phi3MLP), hand-written for sm_100a:
//
//   * operands staged by TMA (cp.async.bulk.tensor, 128-byte swizzle) through a multi-stage
//     shared-memory ring guarded by full/empty mbarriers,
//   * tcgen05.mma (kind::f16, M=128, N=BLOCK_N, K=16) issued by ONE thread, fp32 accumulators in
//     TMEM, double-buffered so the epilogue of tile i overlaps the MMAs of tile i+1,
//   * persistent CTAs (one per SM) walking the tile list m-fastest so a W tile is shared
//     through L2 by the CTAs that run concurrently,
//   * epilogues fused into the TMEM read-back: plain store, +residual (in place on the
//     residual stream), and SwiGLU over block-interleaved gate/up columns.
//
// Warp roles (192 threads): warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer,
// warps 2..5 = epilogue (TMEM lane quadrant = warp_idx % 4).
#include "common.cuh"
#include "vgpt_internal.h"

#include <cuda.h>
#include <cstdlib>

namespace vgpt {

constexpr int kBlockM = 128;
constexpr int kBlockK = 64;   // 64 bf16 = 128 bytes = one swizzle-128B row
constexpr int kUmmaK = 16;
constexpr int kGemmThreads = 192;

template <int BN>
struct GemmCfg {
  static constexpr int kStages = (BN == 256) ? 4 : (BN == 192 ? 5 : 6);
  static constexpr int kABytes = kBlockM * kBlockK * 2;
  static constexpr int kBBytes = BN * kBlockK * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kTmemCols = (2 * BN <= 256) ? 256 : 512;  // two accumulator stages
  static constexpr int kBarBytes = 256;
  static constexpr int kSmemBytes = kStages * kStageBytes + kBarBytes + 1024;  // +1024: alignment
};

enum GemmEpilogue : int { kEpiStore = 0, kEpiResidual = 1, kEpiSwiGLU = 2 };
// defaults used when the caller passes block_n = 0 / cta_pair = -1 (tuned on B200, profiles/)
constexpr int kDefaultCtaPair = 1;

// Tile width for the CTA-pair kernel: fewest (waves x tile width), with the narrower tile charged
// for its higher L2 -> SMEM fill rate per MMA cycle (profiles/r01c_gemm_sweep_pair.txt: qkv and
// gate_up prefer 256, the N = 3072 projections prefer 192 at M = 2064).
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

// One 32-column chunk of one accumulator row -> global memory.
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
        // reference: o_proj/down_proj output is rounded to bf16, then added to the bf16
        // residual stream and rounded again (Phi3DecoderLayer.forward, transformers 4.47.1)
        uint32_t rv = (&r[i].x)[j];
        a = rbf(a) + bf16lo(rv);
        b = rbf(b) + bf16hi(rv);
      }
      w[j] = pack_bf16x2(a, b);
    }
    op[i] = make_uint4(w[0], w[1], w[2], w[3]);
  }
}

template <int BN, int EPI>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_bf16_tcgen05_kernel(const __grid_constant__ CUtensorMap tmap_a,
                         const __grid_constant__ CUtensorMap tmap_b, __nv_bfloat16* __restrict__ C,
                         const __nv_bfloat16* __restrict__ R, int M, int N, int K, int ldc, int flags) {
  using Cfg = GemmCfg<BN>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;   // swizzle-128B atoms: 1 KB aligned
  const uint32_t bar_base = smem_base + Cfg::kStages * Cfg::kStageBytes;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (Cfg::kStages + s); };
  auto tmem_full_bar = [&](int s) { return bar_base + 8u * (2 * Cfg::kStages + s); };
  auto tmem_empty_bar = [&](int s) { return bar_base + 8u * (2 * Cfg::kStages + 2 + s); };
  const uint32_t tmem_slot = bar_base + 8u * (2 * Cfg::kStages + 4);
  uint32_t* tmem_slot_ptr =
      reinterpret_cast<uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int m_tiles = (M + kBlockM - 1) / kBlockM;
  const int n_tiles = (N + BN - 1) / BN;
  const int num_tiles = m_tiles * n_tiles;
  const int k_blocks = K / kBlockK;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_b);
    for (int s = 0; s < Cfg::kStages; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(tmem_full_bar(s), 1);
      mbar_init(tmem_empty_bar(s), 128);
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<Cfg::kTmemCols>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  if (warp == 0) {
    // ================================ TMA producer ================================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int m0 = (tile % m_tiles) * kBlockM;
        const int n0 = (tile / m_tiles) * BN;
        for (int kb = 0; kb < k_blocks; ++kb) {
          mbar_wait(empty_bar(stage), phase ^ 1);
          const uint32_t sa = smem_base + stage * Cfg::kStageBytes;
          const uint32_t sb = sa + Cfg::kABytes;
          if (flags & 1) {                 // debug: MMA-bound ceiling, operands not refreshed
            mbar_arrive(full_bar(stage));
          } else {
            mbar_arrive_expect_tx(full_bar(stage), Cfg::kStageBytes);
            tma_load_2d(sa, &tmap_a, full_bar(stage), kb * kBlockK, m0);
            tma_load_2d(sb, &tmap_b, full_bar(stage), kb * kBlockK, n0);
          }
          if (++stage == Cfg::kStages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ================================ MMA issuer ================================
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_bf16(kBlockM, BN);
      int stage = 0;
      uint32_t phase = 0;
      int local = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++local) {
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
          // K-major, 128B swizzle: rows 128 B apart, 8-row groups 1024 B apart (SBO); advancing
          // 16 elements along K inside the swizzle atom = +32 B on the start address.
          const uint64_t da = make_smem_desc(sa, 16, 1024, kLayoutSW128);
          const uint64_t db = make_smem_desc(sb, 16, 1024, kLayoutSW128);
#pragma unroll
          for (int k = 0; k < kBlockK / kUmmaK; ++k) {
            umma_f16_ss(tmem_d, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), idesc,
                        (kb > 0 || k > 0) ? 1u : 0u);
          }
          umma_commit(empty_bar(stage));   // frees the smem slot once these MMAs have read it
          if (++stage == Cfg::kStages) { stage = 0; phase ^= 1; }
        }
        umma_commit(tmem_full_bar(as));    // accumulator complete -> epilogue
      }
    }
  } else {
    // ================================ epilogue ================================
    const int quad = warp & 3;             // TMEM lanes [32*quad, 32*quad+32)
    int local = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++local) {
      const int as = local & 1;
      const uint32_t aphase = (local >> 1) & 1;
      const int m0 = (tile % m_tiles) * kBlockM;
      const int n0 = (tile / m_tiles) * BN;
      const int row = m0 + quad * 32 + lane;
      mbar_wait(tmem_full_bar(as), aphase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + as * BN;
      if constexpr (EPI == kEpiSwiGLU) {
        // packed W rows: per 64 columns [gate x32 | up x32] of 32 consecutive outputs
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
              // reference (Phi3MLP): gate_up rounded to bf16, silu(gate) rounded, product rounded
              float gv = rbf(__uint_as_float(g[i]));
              float uv = rbf(__uint_as_float(u[i]));
              g[i] = __float_as_uint(uv * rbf(silu_f(gv)));
            }
            store_chunk<kEpiStore>(g, crow + c * 32, nullptr);
          }
        }
      } else {
        __nv_bfloat16* crow = C + (size_t)row * ldc + n0;
        const __nv_bfloat16* rrow = (EPI == kEpiResidual) ? R + (size_t)row * ldc + n0 : nullptr;
#pragma unroll 1
        for (int c = 0; c < BN / 32; ++c) {
          uint32_t acc[32];
          tmem_ld_32x32b_x32(taddr + c * 32, acc);
          tmem_ld_wait();
          if (row < M && n0 + c * 32 < N) store_chunk<EPI>(acc, crow + c * 32, rrow + c * 32);
        }
      }
      tc_fence_before();
      mbar_arrive(tmem_empty_bar(as));
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<Cfg::kTmemCols>(tmem_base);
  }
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
int debug_gemm_flags() {   // VGPT_DEBUG_GEMM_FLAGS=1: skip the TMA loads (profiling experiments only)
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("VGPT_DEBUG_GEMM_FLAGS");
    v = e ? atoi(e) : 0;
  }
  return v;
}

static int make_tmap_2d(CUtensorMap* map, const void* base, uint64_t inner, uint64_t outer,
                        uint64_t pitch_bytes, uint32_t box_inner, uint32_t box_outer,
                        CUtensorMapSwizzle swz) {
  cuuint64_t dims[2] = {inner, outer};
  cuuint64_t strides[1] = {pitch_bytes};
  cuuint32_t box[2] = {box_inner, box_outer};
  cuuint32_t estr[2] = {1, 1};
  return encode_tensor_map(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims,
                           strides, box, estr, swz);
}

template <int BN, int EPI>
static int launch_gemm(const void* A, const void* W, void* C, const void* R, int M, int N, int K,
                       int lda, int ldc, int num_sms, cudaStream_t stream) {
  using Cfg = GemmCfg<BN>;
  CUtensorMap ta, tb;
  int rc = make_tmap_2d(&ta, A, (uint64_t)K, (uint64_t)M, (uint64_t)lda * 2, kBlockK, kBlockM,
                        CU_TENSOR_MAP_SWIZZLE_128B);
  if (rc) return rc;
  rc = make_tmap_2d(&tb, W, (uint64_t)K, (uint64_t)N, (uint64_t)K * 2, kBlockK, BN,
                    CU_TENSOR_MAP_SWIZZLE_128B);
  if (rc) return rc;
  auto kern = gemm_bf16_tcgen05_kernel<BN, EPI>;
  static bool attr_set = false;
  if (!attr_set) {
    VGPT_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         Cfg::kSmemBytes));
    attr_set = true;
  }
  const int tiles = ((M + kBlockM - 1) / kBlockM) * ((N + BN - 1) / BN);
  const int grid = tiles < num_sms ? tiles : num_sms;
  kern<<<grid, kGemmThreads, Cfg::kSmemBytes, stream>>>(
      ta, tb, static_cast<__nv_bfloat16*>(C), static_cast<const __nv_bfloat16*>(R), M, N, K, ldc,
      debug_gemm_flags());
  VGPT_CHECK_LAUNCH();
  return 0;
}

int gemm_bf16(const void* A, const void* W, void* C, const void* R, int M, int N, int K, int lda,
              int ldc, int epilogue, int block_n, int cta_pair, cudaStream_t stream) {
  VGPT_CHECK_ARG(A && W && C, "vgpt_gemm_bf16: null pointer");
  VGPT_CHECK_ARG(M > 0 && N > 0 && K > 0, "vgpt_gemm_bf16: empty problem M=%d N=%d K=%d", M, N, K);
  VGPT_CHECK_ARG(K % kBlockK == 0, "vgpt_gemm_bf16: K=%d must be a multiple of %d", K, kBlockK);
  VGPT_CHECK_ARG(N % 64 == 0, "vgpt_gemm_bf16: N=%d must be a multiple of 64", N);
  VGPT_CHECK_ARG(lda >= K && lda % 8 == 0, "vgpt_gemm_bf16: lda=%d invalid", lda);
  VGPT_CHECK_ARG(ldc % 8 == 0, "vgpt_gemm_bf16: ldc=%d must be a multiple of 8", ldc);
  VGPT_CHECK_ARG(((uintptr_t)A & 15) == 0 && ((uintptr_t)W & 15) == 0 && ((uintptr_t)C & 15) == 0,
                 "vgpt_gemm_bf16: pointers must be 16-byte aligned");
  VGPT_CHECK_ARG(epilogue >= 0 && epilogue <= 2, "vgpt_gemm_bf16: unknown epilogue %d", epilogue);
  VGPT_CHECK_ARG(epilogue != kEpiResidual || R, "vgpt_gemm_bf16: residual epilogue needs R");
  const bool auto_pair = cta_pair < 0;
  if (cta_pair < 0) cta_pair = kDefaultCtaPair;
  // Rows that do not fill a 256-row tile (M % 256 <= 128: the 16 tag / time-slot rows of M = 2064 at cfg2, the
  // remainders of sequence-parallel shards) are computed by swapped-operand TAIL TILES inside the same persistent
  // launch (gemm2_tcgen05.cu, TailArgs) instead of a whole extra row of mostly empty 256-row tiles.  Bit-exact against
  // plain tiles on hardware, but as measured (profiles/r02b_gemm_sweep_fused_tail.txt) a tail tile costs 1.3 - 1.6 main
  // tiles instead of the ~0.3 its bytes and MMAs are worth: its k-loop has 156 cycles of MMA work per stage, the ring
  // holds 6 stages, and a stage's round trip (MMA completion -> commit -> TMA issue -> data) is several thousand cycles,
  // so it runs latency-bound.  Opt-in (VGPT_GEMM_FUSED_TAIL=1, or cta_pair == 3 with any block_n) until that is fixed.
  static const bool fused_tail_on = [] { const char* e = getenv("VGPT_GEMM_FUSED_TAIL"); return e && e[0] == '1'; }();
  const int tail = M % 256;
  if ((cta_pair == 3 || (auto_pair && cta_pair == 1 && fused_tail_on && block_n == 0)) && tail > 0 && tail <= 128 &&
      N % 256 == 0 && N / 256 <= 64) {
    const int sms = device_sm_count();
    const int rows_main = M - tail;
    const int bn = block_n ? block_n : pick_pair_block_n(rows_main > 0 ? rows_main : 256, N, sms);
    VGPT_CHECK_ARG(bn == 128 || bn == 192 || bn == 256, "vgpt_gemm_bf16: CTA-pair block_n must be 128, 192 or 256");
    return gemm_bf16_pair(A, W, C, R, rows_main, N, K, lda, ldc, epilogue, bn, stream, tail);
  }
  if (cta_pair == 3) cta_pair = 1;
  // Earlier form of the same idea, kept for A/B timing only (cta_pair == 2): the tail rows in a SEPARATE launch of the
  // swapped-operand kernel -- it streams W a second time.
  if (cta_pair == 2 && tail > 0 && tail <= 32 && M > 256 && N % 256 == 0 && block_n == 0) {
    const int sms = device_sm_count();
    const int rows_main = M - tail;
    int rc = gemm_bf16_pair(A, W, C, R, rows_main, N, K, lda, ldc, epilogue, pick_pair_block_n(rows_main, N, sms), stream);
    if (rc) return rc;
    const __nv_bfloat16* a_tail = static_cast<const __nv_bfloat16*>(A) + (size_t)rows_main * lda;
    __nv_bfloat16* c_tail = static_cast<__nv_bfloat16*>(C) + (size_t)rows_main * ldc;
    const __nv_bfloat16* r_tail = R ? static_cast<const __nv_bfloat16*>(R) + (size_t)rows_main * ldc : nullptr;
    return gemm_bf16_skinny(a_tail, W, c_tail, r_tail, tail, N, K, lda, ldc, epilogue, stream);
  }
  if (cta_pair == 2) cta_pair = 1;
  if (block_n == 0) block_n = cta_pair ? pick_pair_block_n(M, N, device_sm_count()) : ((N % 256 == 0) ? 256 : 128);
  if (cta_pair) {
    VGPT_CHECK_ARG(block_n == 128 || block_n == 192 || block_n == 256,
                   "vgpt_gemm_bf16: CTA-pair block_n must be 128, 192 or 256");
    return gemm_bf16_pair(A, W, C, R, M, N, K, lda, ldc, epilogue, block_n, stream);
  }
  VGPT_CHECK_ARG((block_n == 128 || block_n == 192 || block_n == 256),
                 "vgpt_gemm_bf16: block_n must be 128, 192 or 256");
  const int sms = device_sm_count();
#define VGPT_GEMM_CASE(BN_, EPI_)                                                         \
  if (block_n == BN_ && epilogue == EPI_)                                                  \
    return launch_gemm<BN_, EPI_>(A, W, C, R, M, N, K, lda, ldc, sms, stream);
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
