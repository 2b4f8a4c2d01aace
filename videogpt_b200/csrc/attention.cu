// Clip-block-causal flash attention over a paged KV cache (replaces the dense-mask SDPA call of
// the reference, LVM/transform/sdpa_transform.py:152-166, and the [B,L,L] additive mask built in
// OmniGen/transformer.py:128-145).
//
// The mask is never materialised.  Every token carries an integer CODE such that
//     allowed(q, k)  <=>  code_q >= code_k
// (for the frame-block layout code = 4 * min(frame, n_ctx) + rank, see processor.py of this
// package; SURVEY.md 8(a) a4 closed form).  KV tiles are classified from precomputed per-tile
// (min, max) codes: fully masked tiles are skipped without being loaded, fully visible tiles
// skip the predicate, boundary tiles evaluate one integer compare per score.
//
// K/V live in a paged pool [page][H][page_tokens=128][D] (post-RoPE K, as the reference caches
// it: sdpa_transform.py:53-57); a per-sequence page table maps logical pages to pool pages.
//
// This is the round-1 kernel: FA2-style, mma.sync m16n8k16 (bf16 in, fp32 accumulate), 128-row
// Q tile per CTA (8 warps x 16 rows), 64-key tiles double-buffered with cp.async.  A tcgen05 /
// TMEM variant replaces it once parity is green (kernels).
#include "common.cuh"
#include "vgpt_internal.h"

namespace vgpt {

constexpr int kAttnBM = 128;
constexpr int kAttnBN = 64;
constexpr int kAttnThreads = 256;
constexpr int kPageTokens = 128;

struct AttnSeq {       // one entry per sequence, device resident (mirrors VgptAttnSeq)
  int32_t q_row0;      // first row of this sequence's queries in q / out / q_code
  int32_t n_q;         // number of query rows
  int32_t kv_len;      // number of valid keys (logical positions [0, kv_len))
  int32_t reserved;
};

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, bool valid) {
  const int sz = valid ? 16 : 0;   // src-size 0 -> zero fill
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(sz)
               : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2,
                                        uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
               : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2,
                                          uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
               : "r"(addr));
}
__device__ __forceinline__ void mma_bf16(float (&c)[4], const uint32_t (&a)[4], uint32_t b0,
                                         uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, "
      "{%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

template <int D>
struct AttnSmem {
  static constexpr int kPitch = D * 2 + 16;                 // bytes per row; +16 B: conflict-free ldmatrix
  static constexpr int kTileBytes = kAttnBN * kPitch;
  static constexpr int kQBytes = kAttnBM * kPitch;
  static constexpr int kCodeBytes = 2 * kAttnBN * 4;
  // [Q][K0][K1][V0][V1][kcode0][kcode1]
  static constexpr int kTotal = kQBytes + 4 * kTileBytes + kCodeBytes;
};

template <int D>
__global__ void __launch_bounds__(kAttnThreads, 1)
attn_clip_causal_kernel(const __nv_bfloat16* __restrict__ q, int q_ld, __nv_bfloat16* __restrict__ out,
                        int out_ld, const __nv_bfloat16* __restrict__ k_pool,
                        const __nv_bfloat16* __restrict__ v_pool, const int32_t* __restrict__ page_table,
                        int max_pages, const AttnSeq* __restrict__ seqs,
                        const int32_t* __restrict__ q_code, const int32_t* __restrict__ k_code,
                        const int32_t* __restrict__ k_tile_minmax, int max_k_tiles, int H,
                        float scale_log2) {
  using S = AttnSmem<D>;
  constexpr int kChunks = D / 8;           // 16-byte chunks per row
  constexpr int kKSteps = D / 16;          // k-steps of QK^T
  constexpr int kDTiles = D / 8;           // n-tiles (of 8) of the output
  extern __shared__ __align__(16) uint8_t smem[];
  const uint32_t s_q = smem_u32(smem);
  const uint32_t s_k = s_q + S::kQBytes;
  const uint32_t s_v = s_k + 2 * S::kTileBytes;
  int32_t* s_code = reinterpret_cast<int32_t*>(smem + S::kQBytes + 4 * S::kTileBytes);

  const int seq_id = blockIdx.z, head = blockIdx.y;
  const AttnSeq sq = seqs[seq_id];
  const int q0 = blockIdx.x * kAttnBM;
  if (q0 >= sq.n_q) return;
  const int rows_here = min(kAttnBM, sq.n_q - q0);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, t4 = lane & 3;

  // ---- Q tile -> smem (async), query codes -> registers ------------------------------------
  for (int i = tid; i < kAttnBM * kChunks; i += kAttnThreads) {
    const int r = i / kChunks, c = i % kChunks;
    const bool ok = r < rows_here;
    const __nv_bfloat16* src =
        q + (size_t)(sq.q_row0 + q0 + (ok ? r : 0)) * q_ld + head * D + c * 8;
    cp_async16(s_q + r * S::kPitch + c * 16, src, ok);
  }
  cp_async_commit();

  const int r_lo = warp * 16 + g, r_hi = r_lo + 8;        // this thread's two rows in the tile
  const int INT_MAXV = 0x7fffffff, INT_MINV = (int)0x80000000;
  const int qc_lo = (r_lo < rows_here) ? q_code[sq.q_row0 + q0 + r_lo] : INT_MINV;
  const int qc_hi = (r_hi < rows_here) ? q_code[sq.q_row0 + q0 + r_hi] : INT_MINV;
  // CTA-wide min / max of the real rows' codes (tile classification)
  __shared__ int s_qmin, s_qmax;
  if (tid == 0) { s_qmin = INT_MAXV; s_qmax = INT_MINV; }
  __syncthreads();
  {
    int mn = min(r_lo < rows_here ? qc_lo : INT_MAXV, r_hi < rows_here ? qc_hi : INT_MAXV);
    int mx = max(qc_lo, qc_hi);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      mn = min(mn, __shfl_xor_sync(0xffffffffu, mn, o));
      mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    }
    if (lane == 0) { atomicMin(&s_qmin, mn); atomicMax(&s_qmax, mx); }
  }
  __syncthreads();
  const int q_min = s_qmin, q_max = s_qmax;

  const int n_kt = (sq.kv_len + kAttnBN - 1) / kAttnBN;
  const int32_t* mm = k_tile_minmax + (size_t)seq_id * max_k_tiles * 2;
  const int32_t* kc = k_code + (size_t)seq_id * max_pages * kPageTokens;
  const int32_t* pt = page_table + (size_t)seq_id * max_pages;

  auto next_tile = [&](int kt) {            // first tile >= kt that is not fully masked
    while (kt < n_kt && mm[2 * kt] > q_max) ++kt;
    return kt;
  };
  auto load_tile = [&](int kt, int buf) {
    const int k0 = kt * kAttnBN;
    const int page = pt[k0 / kPageTokens];
    const int off0 = k0 % kPageTokens;
    const size_t base = (((size_t)page * H + head) * kPageTokens + off0) * D;
    for (int i = tid; i < kAttnBN * kChunks; i += kAttnThreads) {
      const int r = i / kChunks, c = i % kChunks;
      const bool ok = k0 + r < sq.kv_len;
      const size_t o = base + (size_t)(ok ? r : 0) * D + c * 8;
      cp_async16(s_k + buf * S::kTileBytes + r * S::kPitch + c * 16, k_pool + o, ok);
      cp_async16(s_v + buf * S::kTileBytes + r * S::kPitch + c * 16, v_pool + o, ok);
    }
    if (tid < kAttnBN) s_code[buf * kAttnBN + tid] = (k0 + tid < sq.kv_len) ? kc[k0 + tid] : INT_MAXV;
  };

  int kt = next_tile(0);
  if (kt < n_kt) load_tile(kt, 0);
  cp_async_commit();

  float o_acc[kDTiles][4];
#pragma unroll
  for (int i = 0; i < kDTiles; ++i) { o_acc[i][0] = o_acc[i][1] = o_acc[i][2] = o_acc[i][3] = 0.f; }
  float m_lo = -INFINITY, m_hi = -INFINITY, l_lo = 0.f, l_hi = 0.f;
  uint32_t qf[kKSteps][4];
  bool q_loaded = false;
  int buf = 0;

  while (kt < n_kt) {
    const int kt_next = next_tile(kt + 1);
    cp_async_wait<0>();
    __syncthreads();                       // tile kt (and Q) landed; previous tile's reads done
    if (kt_next < n_kt) load_tile(kt_next, buf ^ 1);
    cp_async_commit();

    if (!q_loaded) {
      q_loaded = true;
#pragma unroll
      for (int ks = 0; ks < kKSteps; ++ks) {
        const int r = warp * 16 + (lane & 15);
        const int c = ks * 2 + (lane >> 4);
        ldsm_x4(s_q + r * S::kPitch + c * 16, qf[ks][0], qf[ks][1], qf[ks][2], qf[ks][3]);
      }
    }

    // ---- S = Q K^T ----------------------------------------------------------------------
    float s_acc[kAttnBN / 8][4];
#pragma unroll
    for (int i = 0; i < kAttnBN / 8; ++i) { s_acc[i][0] = s_acc[i][1] = s_acc[i][2] = s_acc[i][3] = 0.f; }
    const uint32_t kb = s_k + buf * S::kTileBytes;
#pragma unroll
    for (int ks = 0; ks < kKSteps; ++ks) {
#pragma unroll
      for (int np = 0; np < kAttnBN / 16; ++np) {       // pairs of 8-key n-tiles
        // matrices: (keys np*16+0..7, d lo), (keys 0..7, d hi), (keys 8..15, d lo), (keys 8..15, d hi)
        const int r = np * 16 + (lane & 7) + ((lane >> 4) << 3);
        const int c = ks * 2 + ((lane >> 3) & 1);
        uint32_t b0, b1, b2, b3;
        ldsm_x4(kb + r * S::kPitch + c * 16, b0, b1, b2, b3);
        mma_bf16(s_acc[2 * np], qf[ks], b0, b1);
        mma_bf16(s_acc[2 * np + 1], qf[ks], b2, b3);
      }
    }

    // ---- mask (boundary tiles only) + online softmax ------------------------------------
    const bool need_mask = mm[2 * kt + 1] > q_min || (kt + 1) * kAttnBN > sq.kv_len;
    const int32_t* codes = s_code + buf * kAttnBN;
    float mx_lo = -INFINITY, mx_hi = -INFINITY;
#pragma unroll
    for (int nt = 0; nt < kAttnBN / 8; ++nt) {
      if (need_mask) {
        const int c0 = codes[nt * 8 + t4 * 2], c1 = codes[nt * 8 + t4 * 2 + 1];
        if (qc_lo < c0) s_acc[nt][0] = -INFINITY;
        if (qc_lo < c1) s_acc[nt][1] = -INFINITY;
        if (qc_hi < c0) s_acc[nt][2] = -INFINITY;
        if (qc_hi < c1) s_acc[nt][3] = -INFINITY;
      }
      mx_lo = fmaxf(mx_lo, fmaxf(s_acc[nt][0], s_acc[nt][1]));
      mx_hi = fmaxf(mx_hi, fmaxf(s_acc[nt][2], s_acc[nt][3]));
    }
    mx_lo = fmaxf(mx_lo, __shfl_xor_sync(0xffffffffu, mx_lo, 1));
    mx_lo = fmaxf(mx_lo, __shfl_xor_sync(0xffffffffu, mx_lo, 2));
    mx_hi = fmaxf(mx_hi, __shfl_xor_sync(0xffffffffu, mx_hi, 1));
    mx_hi = fmaxf(mx_hi, __shfl_xor_sync(0xffffffffu, mx_hi, 2));
    const float mn_lo = fmaxf(m_lo, mx_lo), mn_hi = fmaxf(m_hi, mx_hi);
    // rows with nothing visible so far keep m = -inf; use 0 as the subtrahend to avoid NaN
    const float sub_lo = (mn_lo == -INFINITY) ? 0.f : mn_lo * scale_log2;
    const float sub_hi = (mn_hi == -INFINITY) ? 0.f : mn_hi * scale_log2;
    const float a_lo = (m_lo == -INFINITY) ? 0.f : exp2f(m_lo * scale_log2 - sub_lo);
    const float a_hi = (m_hi == -INFINITY) ? 0.f : exp2f(m_hi * scale_log2 - sub_hi);
    m_lo = mn_lo; m_hi = mn_hi;
    float sum_lo = 0.f, sum_hi = 0.f;
    uint32_t pf[kAttnBN / 16][4];
#pragma unroll
    for (int nt = 0; nt < kAttnBN / 8; ++nt) {
      const float p0 = exp2f(s_acc[nt][0] * scale_log2 - sub_lo);
      const float p1 = exp2f(s_acc[nt][1] * scale_log2 - sub_lo);
      const float p2 = exp2f(s_acc[nt][2] * scale_log2 - sub_hi);
      const float p3 = exp2f(s_acc[nt][3] * scale_log2 - sub_hi);
      sum_lo += p0 + p1;
      sum_hi += p2 + p3;
      pf[nt >> 1][(nt & 1) * 2] = pack_bf16x2(p0, p1);
      pf[nt >> 1][(nt & 1) * 2 + 1] = pack_bf16x2(p2, p3);
    }
    l_lo = l_lo * a_lo + sum_lo;
    l_hi = l_hi * a_hi + sum_hi;
#pragma unroll
    for (int i = 0; i < kDTiles; ++i) {
      o_acc[i][0] *= a_lo; o_acc[i][1] *= a_lo; o_acc[i][2] *= a_hi; o_acc[i][3] *= a_hi;
    }

    // ---- O += P V -----------------------------------------------------------------------
    const uint32_t vb = s_v + buf * S::kTileBytes;
#pragma unroll
    for (int kk = 0; kk < kAttnBN / 16; ++kk) {          // 16 keys per k-step
#pragma unroll
      for (int dp = 0; dp < kDTiles / 2; ++dp) {         // pairs of 8-wide d tiles
        // transposed load: matrices (keys 0..7, d tile 2dp), (keys 8..15, 2dp), (0..7, 2dp+1), (8..15, 2dp+1)
        const int r = kk * 16 + (lane & 7) + (((lane >> 3) & 1) << 3);
        const int c = dp * 2 + (lane >> 4);
        uint32_t b0, b1, b2, b3;
        ldsm_x4_t(vb + r * S::kPitch + c * 16, b0, b1, b2, b3);
        mma_bf16(o_acc[2 * dp], pf[kk], b0, b1);
        mma_bf16(o_acc[2 * dp + 1], pf[kk], b2, b3);
      }
    }
    kt = kt_next;
    buf ^= 1;
  }
  cp_async_wait<0>();

  // ---- normalise and store ---------------------------------------------------------------
  l_lo += __shfl_xor_sync(0xffffffffu, l_lo, 1);
  l_lo += __shfl_xor_sync(0xffffffffu, l_lo, 2);
  l_hi += __shfl_xor_sync(0xffffffffu, l_hi, 1);
  l_hi += __shfl_xor_sync(0xffffffffu, l_hi, 2);
  const float inv_lo = l_lo > 0.f ? 1.f / l_lo : 0.f;
  const float inv_hi = l_hi > 0.f ? 1.f / l_hi : 0.f;
  __nv_bfloat16* o_lo = out + (size_t)(sq.q_row0 + q0 + r_lo) * out_ld + head * D + t4 * 2;
  __nv_bfloat16* o_hi = out + (size_t)(sq.q_row0 + q0 + r_hi) * out_ld + head * D + t4 * 2;
#pragma unroll
  for (int i = 0; i < kDTiles; ++i) {
    if (r_lo < rows_here)
      *reinterpret_cast<uint32_t*>(o_lo + i * 8) = pack_bf16x2(o_acc[i][0] * inv_lo, o_acc[i][1] * inv_lo);
    if (r_hi < rows_here)
      *reinterpret_cast<uint32_t*>(o_hi + i * 8) = pack_bf16x2(o_acc[i][2] * inv_hi, o_acc[i][3] * inv_hi);
  }
}

template <int D>
static int launch_attn(const void* q, int q_ld, void* out, int out_ld, const void* k_pool,
                       const void* v_pool, const int32_t* page_table, int max_pages, const void* seqs,
                       int num_seqs, int max_q_tiles, const int32_t* q_code, const int32_t* k_code,
                       const int32_t* k_tile_minmax, int max_k_tiles, int H, float scale,
                       cudaStream_t s) {
  using S = AttnSmem<D>;
  auto kern = attn_clip_causal_kernel<D>;
  // per launch: the attribute is per device, and a process may drive several devices (cheap, capture-safe)
  VGPT_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, S::kTotal));
  dim3 grid(max_q_tiles, H, num_seqs);
  kern<<<grid, kAttnThreads, S::kTotal, s>>>(
      (const __nv_bfloat16*)q, q_ld, (__nv_bfloat16*)out, out_ld, (const __nv_bfloat16*)k_pool,
      (const __nv_bfloat16*)v_pool, page_table, max_pages, (const AttnSeq*)seqs, q_code, k_code,
      k_tile_minmax, max_k_tiles, H, scale * 1.4426950408889634f);
  VGPT_CHECK_LAUNCH();
  return 0;
}

int attn_clip_causal(const void* q, int q_ld, void* out, int out_ld, const void* k_pool,
                     const void* v_pool, const int32_t* page_table, int max_pages, const void* seqs,
                     int num_seqs, int max_q_rows, const int32_t* q_code, const int32_t* k_code,
                     const int32_t* k_tile_minmax, int max_k_tiles, int H, int D, float scale,
                     cudaStream_t s) {
  VGPT_CHECK_ARG(q && out && k_pool && v_pool && page_table && seqs && q_code && k_code && k_tile_minmax,
                 "vgpt_attn_clip_causal: null pointer");
  VGPT_CHECK_ARG(H > 0 && (D == 64 || D == 96 || D == 128), "vgpt_attn_clip_causal: head_dim %d unsupported (64, 96, 128)", D);
  VGPT_CHECK_ARG(q_ld % 8 == 0 && out_ld % 2 == 0 && q_ld >= H * D && out_ld >= H * D,
                 "vgpt_attn_clip_causal: bad leading dimensions q_ld=%d out_ld=%d", q_ld, out_ld);
  VGPT_CHECK_ARG(max_pages > 0 && max_k_tiles >= max_pages * (kPageTokens / kAttnBN),
                 "vgpt_attn_clip_causal: max_k_tiles=%d too small for max_pages=%d", max_k_tiles, max_pages);
  if (num_seqs <= 0 || max_q_rows <= 0) return 0;
  const int q_tiles = (max_q_rows + kAttnBM - 1) / kAttnBM;
  if (D == 64)
    return launch_attn<64>(q, q_ld, out, out_ld, k_pool, v_pool, page_table, max_pages, seqs, num_seqs,
                           q_tiles, q_code, k_code, k_tile_minmax, max_k_tiles, H, scale, s);
  if (D == 96)
    return launch_attn<96>(q, q_ld, out, out_ld, k_pool, v_pool, page_table, max_pages, seqs, num_seqs,
                           q_tiles, q_code, k_code, k_tile_minmax, max_k_tiles, H, scale, s);
  return launch_attn<128>(q, q_ld, out, out_ld, k_pool, v_pool, page_table, max_pages, seqs, num_seqs,
                          q_tiles, q_code, k_code, k_tile_minmax, max_k_tiles, H, scale, s);
}

}  // namespace vgpt
