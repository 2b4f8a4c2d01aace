// Clip-block-causal flash attention on the 5th-gen tensor cores (tcgen05 + TMEM + TMA).
//
// Same contract as attention.cu (codes instead of a mask, paged KV pools, tile classification
// from per-tile code min/max); this is the Blackwell-native kernel:
//
//   * Q tile (128 rows) and K / V tiles (128 keys = one cache page) are moved by TMA into
//     swizzled shared memory; head_dim 96 is handled as three 32-element (64-byte, SW64) K-major
//     chunks, head_dim 64 / 128 as 64-element SW128 chunks.
//   * S = Q K^T is a tcgen05.mma (M=128, N=128) into TMEM, double buffered so the tensor core
//     computes S(j+1) while the softmax warps work on S(j).
//   * Softmax: 8 warps, two threads per query row (TMEM lane == row, 64 key columns each),
//     exp2 domain, code predicate only on boundary tiles.  P is written back to TENSOR MEMORY
//     (bf16 pairs, double buffered) and consumed from there as the A operand of O += P V.
//   * O += P V is a tcgen05.mma with V consumed exactly as it lies in the cache ([key][d] =
//     MN-major B operand), accumulating in TMEM; when a row maximum grows, the owning thread
//     rescales its O row in TMEM (tcgen05.ld / st) before the next P V is issued.
//
// Warp roles (320 threads): warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer,
// warps 2..9 = softmax / correction / epilogue (TMEM lane quadrant = warp_idx % 4; the two warps
// of a quadrant split the key columns and the output columns in halves).
#include "common.cuh"
#include "vgpt_internal.h"

#include <cuda.h>

namespace vgpt {

constexpr int kTcBM = 128;          // queries per CTA
constexpr int kTcBN = 128;          // keys per tile = one KV page
constexpr int kTcThreads = 320;    // TMA warp + MMA warp + 8 softmax warps

struct AttnSeqTc { int32_t q_row0, n_q, kv_len, reserved; };

template <int D>
struct TcCfg {
  static constexpr int kCW = (D == 96) ? 32 : 64;            // elements per swizzled chunk row
  static constexpr int kRowBytes = kCW * 2;                   // 64 (SW64) or 128 (SW128)
  static constexpr uint32_t kLayout = (D == 96) ? kLayoutSW64 : kLayoutSW128;
  static constexpr int kChunks = D / kCW;
  static constexpr int kChunkBytes = 128 * kRowBytes;         // 128 rows per chunk
  static constexpr int kTileBytes = kChunks * kChunkBytes;    // Q, K or V tile = 128 * D * 2
  static constexpr int kStages = (D == 128) ? 2 : 4;          // K/V ring depth (smem budget)
  static constexpr int kSmem = kTileBytes * (1 + 2 * kStages) + 256 + 1024;
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
__device__ __forceinline__ void tmem_st_32x32b_x16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
      "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
template <int D>
__global__ void __launch_bounds__(kTcThreads, 1)
attn_tcgen05_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_k,
                    const __grid_constant__ CUtensorMap tmap_v, __nv_bfloat16* __restrict__ out, int out_ld,
                    const int32_t* __restrict__ page_table, int max_pages,
                    const AttnSeqTc* __restrict__ seqs, const int32_t* __restrict__ q_code,
                    const int32_t* __restrict__ k_code, const int32_t* __restrict__ k_tile_minmax,
                    int max_k_tiles64, int H, float scale_log2) {
  using C = TcCfg<D>;
  constexpr int kTcStages = C::kStages;
  extern __shared__ uint8_t smem_raw[];
  __shared__ int s_qmin, s_qmax;
  __shared__ float s_rowmax[2][2][128];
  __shared__ float s_rowsum[2][128];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gen = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t s_q = base;
  const uint32_t s_kv = s_q + C::kTileBytes;                  // stage s: K at s_kv + 2*s*tile, V right after
  const uint32_t bars = s_kv + 2 * kTcStages * C::kTileBytes;
  const uint32_t bar_q = bars;                                 // Q landed
  auto bar_kv_full = [&](int s) { return bars + 8u * (1 + s); };
  auto bar_kv_empty = [&](int s) { return bars + 8u * (1 + kTcStages + s); };
  auto bar_s_full = [&](int s) { return bars + 8u * (1 + 2 * kTcStages + s); };
  const uint32_t bar_p_full = bars + 8u * (3 + 2 * kTcStages);
  const uint32_t bar_o_full = bars + 8u * (4 + 2 * kTcStages);
  const uint32_t tmem_slot = bars + 8u * (5 + 2 * kTcStages);
  uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(gen + (tmem_slot - base));

  const int seq_id = blockIdx.z, head = blockIdx.y;
  const AttnSeqTc sq = seqs[seq_id];
  const int q0 = blockIdx.x * kTcBM;
  if (q0 >= sq.n_q) return;
  const int rows_here = min(kTcBM, sq.n_q - q0);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  // ---- CTA-wide min / max of the query codes (tile classification) -----------------------------
  if (threadIdx.x == 0) { s_qmin = 0x7fffffff; s_qmax = (int)0x80000000; }
  __syncthreads();
  if (threadIdx.x < kTcBM && (int)threadIdx.x < rows_here) {
    const int c = q_code[sq.q_row0 + q0 + threadIdx.x];
    atomicMin(&s_qmin, c);
    atomicMax(&s_qmax, c);
  }
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_q); tma_prefetch_desc(&tmap_k); tma_prefetch_desc(&tmap_v);
    mbar_init(bar_q, 1);
    for (int s = 0; s < kTcStages; ++s) { mbar_init(bar_kv_full(s), 1); mbar_init(bar_kv_empty(s), 1); }
    mbar_init(bar_s_full(0), 1); mbar_init(bar_s_full(1), 1);
    mbar_init(bar_p_full, 256);
    mbar_init(bar_o_full, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<512>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot_ptr;
  // TMEM columns: S[0] 0..127, S[1] 128..255, O 256..(256+D), P[0] 384..447, P[1] 448..511
  // (P = bf16 probabilities, two per 32-bit column: the A operand of O += P V)
  const uint32_t tmem_s0 = tmem, tmem_o = tmem + 256, tmem_p0 = tmem + 384;
  const int q_min = s_qmin, q_max = s_qmax;

  const int n_kt = (sq.kv_len + kTcBN - 1) / kTcBN;
  const int32_t* mm = k_tile_minmax + (size_t)seq_id * max_k_tiles64 * 2;   // entries per 64 keys
  const int32_t* pt = page_table + (size_t)seq_id * max_pages;
  auto tile_min = [&](int kt) { return min(mm[4 * kt], mm[4 * kt + 2]); };
  auto tile_max = [&](int kt) { return max(mm[4 * kt + 1], mm[4 * kt + 3]); };
  auto next_tile = [&](int kt) {          // first tile >= kt that is not fully masked for this Q tile
    while (kt < n_kt && tile_min(kt) > q_max) ++kt;
    return kt;
  };

  if (warp == 0) {
    // =================================== TMA producer ===================================
    if (lane == 0) {
      mbar_arrive_expect_tx(bar_q, C::kTileBytes);
#pragma unroll
      for (int c = 0; c < C::kChunks; ++c)
        tma_load_2d(s_q + c * C::kChunkBytes, &tmap_q, bar_q, head * D + c * C::kCW, sq.q_row0 + q0);
      int stage = 0; uint32_t phase = 0;
      for (int kt = next_tile(0); kt < n_kt; kt = next_tile(kt + 1)) {
        mbar_wait(bar_kv_empty(stage), phase ^ 1);
        const uint32_t sk = s_kv + 2 * stage * C::kTileBytes, sv = sk + C::kTileBytes;
        const int row = (pt[kt] * H + head) * kTcBN;           // pool viewed as [(page*H + head)*128 + tok][D]
        mbar_arrive_expect_tx(bar_kv_full(stage), 2 * C::kTileBytes);
#pragma unroll
        for (int c = 0; c < C::kChunks; ++c) {
          tma_load_2d(sk + c * C::kChunkBytes, &tmap_k, bar_kv_full(stage), c * C::kCW, row);
          tma_load_2d(sv + c * C::kChunkBytes, &tmap_v, bar_kv_full(stage), c * C::kCW, row);
        }
        if (++stage == kTcStages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    // =================================== MMA issuer ===================================
    if (lane == 0) {
      constexpr uint32_t idesc_s = make_idesc_bf16(128, kTcBN);          // S = Q K^T (both K-major)
      constexpr uint32_t idesc_o = make_idesc_bf16(128, D, 0, 1);        // O = P V   (V MN-major)
      auto issue_s = [&](int stage, int sbuf) {
        const uint32_t sk = s_kv + 2 * stage * C::kTileBytes;
#pragma unroll
        for (int c = 0; c < C::kChunks; ++c) {
#pragma unroll
          for (int ks = 0; ks < C::kCW / 16; ++ks) {
            const uint64_t da = make_smem_desc(s_q + c * C::kChunkBytes + ks * 32, 16, 8 * C::kRowBytes, C::kLayout);
            const uint64_t db = make_smem_desc(sk + c * C::kChunkBytes + ks * 32, 16, 8 * C::kRowBytes, C::kLayout);
            umma_f16_ss(tmem_s0 + sbuf * 128, da, db, idesc_s, (c | ks) ? 1u : 0u);
          }
        }
        umma_commit(bar_s_full(sbuf));
      };
      mbar_wait(bar_q, 0);
      int kt = next_tile(0);
      int stage_s = 0; uint32_t phase_s = 0;      // stage / phase of the next S to issue
      int stage_o = 0;                            // stage of the next P V
      int issued = 0;
      // prologue: S(0), S(1)
      int kt_s = kt;
      for (int i = 0; i < 2 && kt_s < n_kt; ++i) {
        mbar_wait(bar_kv_full(stage_s), phase_s);
        tc_fence_after();
        issue_s(stage_s, issued & 1);
        ++issued;
        if (++stage_s == kTcStages) { stage_s = 0; phase_s ^= 1; }
        kt_s = next_tile(kt_s + 1);
      }
      int j = 0;
      for (; kt < n_kt; kt = next_tile(kt + 1), ++j) {
        mbar_wait(bar_p_full, j & 1);              // P(j) in smem, O rescaled, S(j) consumed
        tc_fence_after();
        const uint32_t sv = s_kv + 2 * stage_o * C::kTileBytes + C::kTileBytes;
#pragma unroll
        for (int ks = 0; ks < kTcBN / 16; ++ks) {
          const uint64_t db = make_smem_desc(sv + ks * 16 * C::kRowBytes, C::kChunkBytes, 8 * C::kRowBytes, C::kLayout);
          umma_f16_ts(tmem_o, tmem_p0 + (j & 1) * 64 + ks * 8, db, idesc_o, (j > 0 || ks > 0) ? 1u : 0u);
        }
        umma_commit(bar_kv_empty(stage_o));        // K(j), V(j) no longer needed
        umma_commit(bar_o_full);                   // O includes tile j; P[j&1] reusable
        if (++stage_o == kTcStages) stage_o = 0;
        if (kt_s < n_kt) {                         // S(j+2) into the buffer softmax(j) just released
          mbar_wait(bar_kv_full(stage_s), phase_s);
          tc_fence_after();
          issue_s(stage_s, issued & 1);
          ++issued;
          if (++stage_s == kTcStages) { stage_s = 0; phase_s ^= 1; }
          kt_s = next_tile(kt_s + 1);
        }
      }
    }
  } else {
    // ========================= softmax / correction / epilogue =========================
    // 8 warps: two per TMEM lane quadrant; the pair splits the 128 key columns of a row (and the
    // D output columns) in halves, so every SM sub-partition has two softmax warps to interleave.
    const int quad = warp & 3;
    const int half = (warp - 2) >> 2;
    const int row = quad * 32 + lane;                       // row of the Q tile == TMEM lane
    const uint32_t lane_addr = (uint32_t)(quad * 32) << 16;
    const int qc = (row < rows_here) ? q_code[sq.q_row0 + q0 + row] : (int)0x80000000;
    const int32_t* kc = k_code + (size_t)seq_id * max_pages * kTcBN + half * 64;
    constexpr int kOUnits = D / 32;                         // 16-column units of O per thread
    const uint32_t o_addr = tmem_o + lane_addr + half * (D / 2);
    float m_run = -INFINITY, l_run = 0.f;
    int j = 0;
    for (int kt = next_tile(0); kt < n_kt; kt = next_tile(kt + 1), ++j) {
      mbar_wait(bar_s_full(j & 1), (j >> 1) & 1);
      tc_fence_after();
      uint32_t s[64];
#pragma unroll
      for (int c = 0; c < 2; ++c) tmem_ld_32x32b_x32(tmem_s0 + lane_addr + (j & 1) * 128 + half * 64 + c * 32,
                                                      *reinterpret_cast<uint32_t(*)[32]>(&s[c * 32]));
      tmem_ld_wait();
      const bool need_mask = tile_max(kt) > q_min || (kt + 1) * kTcBN > sq.kv_len;
      if (need_mask) {
        const int4* kcode = reinterpret_cast<const int4*>(kc + kt * kTcBN);
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const int4 c4 = __ldg(kcode + i);
          if (qc < c4.x) s[4 * i + 0] = 0xff800000u;          // -inf
          if (qc < c4.y) s[4 * i + 1] = 0xff800000u;
          if (qc < c4.z) s[4 * i + 2] = 0xff800000u;
          if (qc < c4.w) s[4 * i + 3] = 0xff800000u;
        }
      }
      float mx = -INFINITY;
#pragma unroll
      for (int i = 0; i < 64; ++i) mx = fmaxf(mx, __uint_as_float(s[i]));
      s_rowmax[j & 1][half][row] = mx;
      asm volatile("bar.sync 1, 256;" ::: "memory");        // the 8 softmax warps only
      const float m_new = fmaxf(m_run, fmaxf(mx, s_rowmax[j & 1][half ^ 1][row]));
      const float sub = (m_new == -INFINITY) ? 0.f : __fmul_rn(m_new, scale_log2);
      // alpha is EXACTLY 1 for a row whose maximum did not move (no FMA contraction of the
      // difference), so the warp-uniform "somebody's maximum grew" rescale below leaves such rows
      // bit-identical whatever rows share their warp: results do not depend on how query rows are
      // grouped into tiles (sequence parallelism relies on this).
      const float alpha = (m_run == -INFINITY) ? 0.f
                          : (m_new == m_run)   ? 1.f
                                               : ex2_approx(__fsub_rn(__fmul_rn(m_run, scale_log2), sub));
      // P half-row -> TMEM buffer j&1 (last read by P V of tile j-2, whose completion every thread
      // observed in iteration j-1), overlapping the P V of tile j-1 on the tensor pipe
      float sum = 0.f;
      uint32_t w[32];
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        const float p0 = ex2_approx(fmaf(__uint_as_float(s[2 * i]), scale_log2, -sub));
        const float p1 = ex2_approx(fmaf(__uint_as_float(s[2 * i + 1]), scale_log2, -sub));
        sum += p0 + p1;
        w[i] = pack_bf16x2(p0, p1);
      }
      tmem_st_32x32b_x32(tmem_p0 + lane_addr + (j & 1) * 64 + half * 32, w);
      l_run = l_run * alpha + sum;
      // O(j-1) must be complete before it is rescaled (and before P V(j) may be issued)
      if (j > 0) {
        mbar_wait(bar_o_full, (j - 1) & 1);
        tc_fence_after();
        if (__any_sync(0xffffffffu, m_new > m_run)) {
#pragma unroll
          for (int u = 0; u < kOUnits; ++u) {
            uint32_t o[16];
            tmem_ld_32x32b_x16(o_addr + u * 16, o);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 16; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
            tmem_st_32x32b_x16(o_addr + u * 16, o);
          }
        }
      }
      m_run = m_new;
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(bar_p_full);
    }
    // ---- epilogue: O / l -> bf16 -> global ------------------------------------------------
    s_rowsum[half][row] = l_run;
    asm volatile("bar.sync 1, 256;" ::: "memory");
    const float l_tot = l_run + s_rowsum[half ^ 1][row];
    if (j > 0) {
      mbar_wait(bar_o_full, (j - 1) & 1);
      tc_fence_after();
    }
    const float inv = l_tot > 0.f ? 1.f / l_tot : 0.f;
    __nv_bfloat16* orow = out + (size_t)(sq.q_row0 + q0 + row) * out_ld + head * D + half * (D / 2);
#pragma unroll
    for (int u = 0; u < kOUnits; ++u) {
      uint32_t o[16];
      if (j > 0) {
        tmem_ld_32x32b_x16(o_addr + u * 16, o);
        tmem_ld_wait();
      } else {
#pragma unroll
        for (int i = 0; i < 16; ++i) o[i] = 0u;
      }
      if (row < rows_here) {
#pragma unroll
        for (int v4 = 0; v4 < 2; ++v4) {
          uint32_t w[4];
#pragma unroll
          for (int h2 = 0; h2 < 4; ++h2)
            w[h2] = pack_bf16x2(__uint_as_float(o[v4 * 8 + h2 * 2]) * inv, __uint_as_float(o[v4 * 8 + h2 * 2 + 1]) * inv);
          reinterpret_cast<uint4*>(orow + u * 16)[v4] = make_uint4(w[0], w[1], w[2], w[3]);
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<512>(tmem);
  }
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
template <int D>
static int launch_attn_tc(const void* q, int q_ld, int q_rows, void* out, int out_ld, const void* k_pool,
                          const void* v_pool, int total_pages, const int32_t* page_table, int max_pages,
                          const void* seqs, int num_seqs, int q_tiles, const int32_t* q_code,
                          const int32_t* k_code, const int32_t* k_tile_minmax, int max_k_tiles64, int H,
                          float scale, cudaStream_t s) {
  using C = TcCfg<D>;
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
  auto kern = attn_tcgen05_kernel<D>;
  static bool attr_set = false;
  if (!attr_set) {
    VGPT_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::kSmem));
    attr_set = true;
  }
  dim3 grid(q_tiles, H, num_seqs);
  kern<<<grid, kTcThreads, C::kSmem, s>>>(tq, tk, tv, (__nv_bfloat16*)out, out_ld, page_table, max_pages,
                                           (const AttnSeqTc*)seqs, q_code, k_code, k_tile_minmax,
                                           max_k_tiles64, H, scale * 1.4426950408889634f);
  VGPT_CHECK_LAUNCH();
  return 0;
}

int attn_clip_causal_tc(const void* q, int q_ld, int q_rows, void* out, int out_ld, const void* k_pool,
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
  if (num_seqs <= 0 || max_q_rows <= 0) return 0;
  const int q_tiles = (max_q_rows + kTcBM - 1) / kTcBM;
#define VGPT_ATTN_CASE(D_)                                                                              \
  if (D == D_)                                                                                           \
    return launch_attn_tc<D_>(q, q_ld, q_rows, out, out_ld, k_pool, v_pool, total_pages, page_table,    \
                              max_pages, seqs, num_seqs, q_tiles, q_code, k_code, k_tile_minmax,        \
                              max_k_tiles, H, scale, s);
  VGPT_ATTN_CASE(64)
  VGPT_ATTN_CASE(96)
  VGPT_ATTN_CASE(128)
#undef VGPT_ATTN_CASE
  return -1;
}

}  // namespace vgpt
