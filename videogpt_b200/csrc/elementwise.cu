// HBM-bound kernels of the next-clip denoising path: sequence assembly (token / time / patch
// embedding), RMSNorm, RoPE + KV-cache append, timestep embedding, small-batch linears, the final
// adaLN layer with unpatchify, the CFG + Euler update, and the code-based mask materialiser.
// All are vectorised (16-byte accesses), one CTA per row unless noted; rounding points follow
// the reference's bf16 PyTorch path (each op's result rounded to bf16), see DESIGN.md.
#include "common.cuh"
#include "vgpt_internal.h"

namespace vgpt {

// ---------------------------------------------------------------------------------------------
// block-wide sum (blockDim.x multiple of 32, <= 1024)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float block_sum(float v, float* red) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  v = warp_sum(v);
  __syncthreads();                 // protect red[] from the previous use
  if (lane == 0) red[warp] = v;
  __syncthreads();
  float t = (lane < nw) ? red[lane] : 0.f;
  return warp_sum(t);
}

__device__ __forceinline__ void unpack8(const uint4& v, float (&f)[8]) {
  f[0] = bf16lo(v.x); f[1] = bf16hi(v.x); f[2] = bf16lo(v.y); f[3] = bf16hi(v.y);
  f[4] = bf16lo(v.z); f[5] = bf16hi(v.z); f[6] = bf16lo(v.w); f[7] = bf16hi(v.w);
}
__device__ __forceinline__ uint4 pack8(const float (&f)[8]) {
  return make_uint4(pack_bf16x2(f[0], f[1]), pack_bf16x2(f[2], f[3]), pack_bf16x2(f[4], f[5]),
                    pack_bf16x2(f[6], f[7]));
}

// ---------------------------------------------------------------------------------------------
// RMSNorm (Phi3RMSNorm, transformers 4.47.1): y = w * bf16(x * rsqrt(mean(x^2) + eps))
// ---------------------------------------------------------------------------------------------
constexpr int kRowThreads = 128;
constexpr int kMaxChunksPerThread = 4;   // hidden <= 128 * 4 * 8 = 4096

__global__ void __launch_bounds__(kRowThreads)
rmsnorm_kernel(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ w,
               __nv_bfloat16* __restrict__ y, int hidden, float eps) {
  __shared__ float red[32];
  const int chunks = hidden >> 3;
  const uint4* xr = reinterpret_cast<const uint4*>(x + (size_t)blockIdx.x * hidden);
  uint4 v[kMaxChunksPerThread];
  float ss = 0.f;
#pragma unroll
  for (int i = 0; i < kMaxChunksPerThread; ++i) {
    const int c = threadIdx.x + i * kRowThreads;
    if (c < chunks) {
      v[i] = xr[c];
      float f[8];
      unpack8(v[i], f);
#pragma unroll
      for (int j = 0; j < 8; ++j) ss += f[j] * f[j];
    }
  }
  const float rstd = rsqrtf(block_sum(ss, red) / (float)hidden + eps);
  uint4* yr = reinterpret_cast<uint4*>(y + (size_t)blockIdx.x * hidden);
  const uint4* wr = reinterpret_cast<const uint4*>(w);
#pragma unroll
  for (int i = 0; i < kMaxChunksPerThread; ++i) {
    const int c = threadIdx.x + i * kRowThreads;
    if (c < chunks) {
      float f[8], g[8];
      unpack8(v[i], f);
      unpack8(wr[c], g);
#pragma unroll
      for (int j = 0; j < 8; ++j) f[j] = g[j] * rbf(f[j] * rstd);
      yr[c] = pack8(f);
    }
  }
}

int rmsnorm(const void* x, const void* w, void* y, int rows, int hidden, float eps, cudaStream_t s) {
  VGPT_CHECK_ARG(x && w && y, "vgpt_rmsnorm: null pointer");
  VGPT_CHECK_ARG(rows >= 0 && hidden > 0 && hidden % 8 == 0 &&
                     hidden <= kRowThreads * kMaxChunksPerThread * 8,
                 "vgpt_rmsnorm: unsupported shape rows=%d hidden=%d", rows, hidden);
  if (rows == 0) return 0;
  rmsnorm_kernel<<<rows, kRowThreads, 0, s>>>((const __nv_bfloat16*)x, (const __nv_bfloat16*)w,
                                              (__nv_bfloat16*)y, hidden, eps);
  VGPT_CHECK_LAUNCH();
  return 0;
}

// ---------------------------------------------------------------------------------------------
// RoPE table: cos/sin of pos * inv_freq in fp32, cast to bf16 (Phi3RotaryEmbedding.forward,
// called at LVM/transform/sdpa_transform.py:52).  Layout [max_pos][D]: cos[0:D/2] | sin[0:D/2].
// ---------------------------------------------------------------------------------------------
__global__ void rope_table_kernel(const float* __restrict__ inv_freq, __nv_bfloat16* __restrict__ tab,
                                  int max_pos, int half) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= max_pos * half) return;
  const int pos = idx / half, i = idx % half;
  const float ang = inv_freq[i] * (float)pos;
  tab[(size_t)pos * 2 * half + i] = __float2bfloat16_rn(cosf(ang));
  tab[(size_t)pos * 2 * half + half + i] = __float2bfloat16_rn(sinf(ang));
}

int rope_table(const float* inv_freq, void* tab, int max_pos, int head_dim, cudaStream_t s) {
  VGPT_CHECK_ARG(inv_freq && tab && max_pos > 0 && head_dim > 0 && head_dim % 2 == 0,
                 "vgpt_rope_table: bad arguments");
  const int n = max_pos * (head_dim / 2);
  rope_table_kernel<<<(n + 255) / 256, 256, 0, s>>>(inv_freq, (__nv_bfloat16*)tab, max_pos,
                                                    head_dim / 2);
  VGPT_CHECK_LAUNCH();
  return 0;
}

// ---------------------------------------------------------------------------------------------
// RoPE (half-split rotation, bf16 rounding of each product and of the sum, as
// apply_rotary_pos_emb does in bf16) applied to q in place and to k on its way into the paged
// KV pool; v is copied into the pool.  Pool layout [page][H][page_tokens][D]; `row_slot` is the
// physical token slot (page * page_tokens + offset) of each row, < 0 = do not cache.
// One CTA per row.  (sdpa_transform.py:39-57: the cache stores post-RoPE K.)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(192)
rope_kv_append_kernel(__nv_bfloat16* __restrict__ qkv, const int32_t* __restrict__ row_pos,
                      const int32_t* __restrict__ row_slot, const __nv_bfloat16* __restrict__ tab,
                      PeerPtrs k_pools, PeerPtrs v_pools, int n_pools, int H, int D, int page_tokens) {
  const int row = blockIdx.x;
  const int half = D >> 1, hc = half >> 3;           // 16-byte chunks per half head
  const int HD = H * D;
  __nv_bfloat16* base = qkv + (size_t)row * 3 * HD;
  const int pos = row_pos[row];
  const int slot = row_slot[row];
  const uint4* cs = reinterpret_cast<const uint4*>(tab + (size_t)pos * D);
  size_t pool_off = 0;
  if (slot >= 0) {
    const int page = slot / page_tokens, off = slot % page_tokens;
    pool_off = ((size_t)page * H * page_tokens + off) * D;   // + h * page_tokens * D per head
  }
  const int items = H * hc;
  for (int it = threadIdx.x; it < 2 * items; it += blockDim.x) {
    const int which = it / items;                    // 0 = q, 1 = k
    const int r = it % items, h = r / hc, c = r % hc;
    __nv_bfloat16* p = base + which * HD + h * D + c * 8;
    float lo[8], hi[8], co[8], si[8], olo[8], ohi[8];
    unpack8(*reinterpret_cast<const uint4*>(p), lo);
    unpack8(*reinterpret_cast<const uint4*>(p + half), hi);
    unpack8(cs[c], co);
    unpack8(cs[hc + c], si);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      olo[j] = rbf(lo[j] * co[j]) + rbf(-hi[j] * si[j]);
      ohi[j] = rbf(hi[j] * co[j]) + rbf(lo[j] * si[j]);
    }
    const uint4 vlo = pack8(olo), vhi = pack8(ohi);
    if (which == 0) {
      *reinterpret_cast<uint4*>(p) = vlo;
      *reinterpret_cast<uint4*>(p + half) = vhi;
    } else if (slot >= 0) {
      const size_t o = pool_off + (size_t)h * page_tokens * D + c * 8;
#pragma unroll
      for (int g = 0; g < kMaxPeers; ++g) {          // own pool + every peer's (NVLink stores)
        if (g < n_pools) {                           // (unrolled: the pointers stay in constant memory)
          __nv_bfloat16* d = static_cast<__nv_bfloat16*>(k_pools.p[g]) + o;
          *reinterpret_cast<uint4*>(d) = vlo;
          *reinterpret_cast<uint4*>(d + half) = vhi;
        }
      }
    }
  }
  if (slot >= 0) {
    const int dc = D >> 3;
    for (int it = threadIdx.x; it < H * dc; it += blockDim.x) {
      const int h = it / dc, c = it % dc;
      const uint4 v = *reinterpret_cast<const uint4*>(base + 2 * HD + h * D + c * 8);
      const size_t o = pool_off + (size_t)h * page_tokens * D + c * 8;
#pragma unroll
      for (int g = 0; g < kMaxPeers; ++g)
        if (g < n_pools) *reinterpret_cast<uint4*>(static_cast<__nv_bfloat16*>(v_pools.p[g]) + o) = v;
    }
  }
}

int rope_kv_append(void* qkv, const int32_t* row_pos, const int32_t* row_slot, const void* tab,
                   void* const* k_pools, void* const* v_pools, int n_pools, int rows, int H, int D,
                   int page_tokens, cudaStream_t s) {
  VGPT_CHECK_ARG(qkv && row_pos && row_slot && tab && k_pools && v_pools,
                 "vgpt_rope_kv_append: null pointer");
  VGPT_CHECK_ARG(n_pools >= 1 && n_pools <= kMaxPeers, "vgpt_rope_kv_append: %d pools (1..%d)", n_pools, kMaxPeers);
  VGPT_CHECK_ARG(H > 0 && D > 0 && D % 16 == 0 && page_tokens > 0,
                 "vgpt_rope_kv_append: unsupported H=%d D=%d page_tokens=%d", H, D, page_tokens);
  PeerPtrs kp, vp;
  for (int i = 0; i < kMaxPeers; ++i) {
    kp.p[i] = i < n_pools ? k_pools[i] : nullptr;
    vp.p[i] = i < n_pools ? v_pools[i] : nullptr;
    VGPT_CHECK_ARG(i >= n_pools || (kp.p[i] && vp.p[i]), "vgpt_rope_kv_append: null pool pointer %d", i);
  }
  if (rows <= 0) return 0;
  rope_kv_append_kernel<<<rows, 192, 0, s>>>((__nv_bfloat16*)qkv, row_pos, row_slot,
                                             (const __nv_bfloat16*)tab, kp, vp, n_pools, H, D, page_tokens);
  VGPT_CHECK_LAUNCH();
  return 0;
}

// ---------------------------------------------------------------------------------------------
// Sequence assembly (LVM/model.py:419-454): every active row of the hidden-state matrix is one of
//   kind 0: embed_tokens[a]                        (tag / pad tokens keep their embedding row)
//   kind 1: time_tokens[a]                         (time slot of generated frame a)
//   kind 2: x_embedder(z[a]) patch b + pos_embed[b]        (noisy latent; PatchEmbedMR 138-154)
//   kind 3: input_x_embedder(ctx[a]) patch b + pos_embed[b] (context latent)
// Conv2d(k=2,s=2) == per-patch linear over the (c,ph,pw)-flattened 16-vector (K=16: FMA, not
// tensor cores); conv output (+bias) rounded to bf16, then + pos_embed rounded to bf16.  One CTA per 8 rows: a thread keeps
// the 16 x 8 weights of its output chunk in registers across the rows (one CTA per row read the 98 KB weight 2064 times).
// ---------------------------------------------------------------------------------------------
constexpr int kEmbedRows = 8;      // rows per CTA: the 98 KB conv weight is read once per 8 rows instead of once per row

__global__ void __launch_bounds__(kRowThreads)
embed_assemble_kernel(__nv_bfloat16* __restrict__ hidden, int rows, int hs, const int32_t* __restrict__ kind,
                      const int32_t* __restrict__ arg_a, const int32_t* __restrict__ arg_b,
                      const __nv_bfloat16* __restrict__ embed_tokens,
                      const __nv_bfloat16* __restrict__ time_tokens,
                      const __nv_bfloat16* __restrict__ z, const __nv_bfloat16* __restrict__ ctx,
                      int C, int lat_h, int lat_w, const __nv_bfloat16* __restrict__ wx,
                      const __nv_bfloat16* __restrict__ bx, const __nv_bfloat16* __restrict__ wc,
                      const __nv_bfloat16* __restrict__ bc, const __nv_bfloat16* __restrict__ pos) {
  __shared__ int s_kind[kEmbedRows], s_a[kEmbedRows], s_b[kEmbedRows];
  __shared__ float s_patch[kEmbedRows][16];
  const int row0 = blockIdx.x * kEmbedRows;
  const int nrows = min(kEmbedRows, rows - row0);
  if ((int)threadIdx.x < nrows) {
    s_kind[threadIdx.x] = kind[row0 + threadIdx.x];
    s_a[threadIdx.x] = arg_a[row0 + threadIdx.x];
    s_b[threadIdx.x] = arg_b[row0 + threadIdx.x];
  }
  __syncthreads();
  const int pw = lat_w >> 1;
  {    // the 16 inputs (c, ph, pw) of every patch row of this CTA
    const int r = threadIdx.x >> 4, t = threadIdx.x & 15;
    if (r < nrows && s_kind[r] >= 2) {
      const int b = s_b[r], py = b / pw, px = b % pw;
      const __nv_bfloat16* lat = (s_kind[r] == 2 ? z : ctx) + (size_t)s_a[r] * C * lat_h * lat_w;
      const int c = t >> 2, ph = (t >> 1) & 1, pq = t & 1;
      s_patch[r][t] = __bfloat162float(lat[((size_t)c * lat_h + 2 * py + ph) * lat_w + 2 * px + pq]);
    }
  }
  __syncthreads();
  bool any2 = false, any3 = false;
  for (int r = 0; r < nrows; ++r) { any2 |= s_kind[r] == 2; any3 |= s_kind[r] == 3; }
  const int chunks = hs >> 3;
  for (int c = threadIdx.x; c < chunks; c += kRowThreads) {
    for (int r = 0; r < nrows; ++r) {      // rows that are plain copies (tag / pad tokens, time slots)
      if (s_kind[r] <= 1) {
        const uint4* src = reinterpret_cast<const uint4*>((s_kind[r] == 0 ? embed_tokens : time_tokens) + (size_t)s_a[r] * hs);
        reinterpret_cast<uint4*>(hidden + (size_t)(row0 + r) * hs)[c] = src[c];
      }
    }
#pragma unroll 1
    for (int kd = 2; kd <= 3; ++kd) {      // patch rows of either embedder: weights of this chunk held in registers
      if (!(kd == 2 ? any2 : any3)) continue;
      const __nv_bfloat16* w = (kd == 2) ? wx : wc;
      float bv[8];
      unpack8(reinterpret_cast<const uint4*>((kd == 2) ? bx : bc)[c], bv);
      uint4 wv[16];
      const uint4* wr = reinterpret_cast<const uint4*>(w + (size_t)c * 8 * 16);
#pragma unroll
      for (int i = 0; i < 16; ++i) wv[i] = wr[i];
      for (int r = 0; r < nrows; ++r) {
        if (s_kind[r] != kd) continue;
        float pv[8], o[8];
        unpack8(reinterpret_cast<const uint4*>(pos + (size_t)s_b[r] * hs)[c], pv);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          float w0[8], w1[8];
          unpack8(wv[2 * j], w0);
          unpack8(wv[2 * j + 1], w1);
          float acc = 0.f;
#pragma unroll
          for (int t = 0; t < 8; ++t) acc = fmaf(w0[t], s_patch[r][t], acc);
#pragma unroll
          for (int t = 0; t < 8; ++t) acc = fmaf(w1[t], s_patch[r][8 + t], acc);
          o[j] = rbf(rbf(acc + bv[j]) + pv[j]);
        }
        reinterpret_cast<uint4*>(hidden + (size_t)(row0 + r) * hs)[c] = pack8(o);
      }
    }
  }
}

int embed_assemble(void* hidden, int rows, int hs, const int32_t* kind, const int32_t* a,
                   const int32_t* b, const void* embed_tokens, const void* time_tokens, const void* z,
                   const void* ctx, int C, int lat_h, int lat_w, const void* wx, const void* bx,
                   const void* wc, const void* bc, const void* pos, cudaStream_t s) {
  VGPT_CHECK_ARG(hidden && kind && a && b, "vgpt_embed_assemble: null pointer");
  VGPT_CHECK_ARG(hs % 8 == 0 && C == 4 && lat_h % 2 == 0 && lat_w % 2 == 0,
                 "vgpt_embed_assemble: unsupported hs=%d C=%d latent %dx%d", hs, C, lat_h, lat_w);
  if (rows <= 0) return 0;
  embed_assemble_kernel<<<(rows + kEmbedRows - 1) / kEmbedRows, kRowThreads, 0, s>>>(
      (__nv_bfloat16*)hidden, rows, hs, kind, a, b, (const __nv_bfloat16*)embed_tokens,
      (const __nv_bfloat16*)time_tokens, (const __nv_bfloat16*)z, (const __nv_bfloat16*)ctx, C,
      lat_h, lat_w, (const __nv_bfloat16*)wx, (const __nv_bfloat16*)bx, (const __nv_bfloat16*)wc,
      (const __nv_bfloat16*)bc, (const __nv_bfloat16*)pos);
  VGPT_CHECK_LAUNCH();
  return 0;
}

// ---------------------------------------------------------------------------------------------
// Timestep sinusoid (TimestepEmbedder.timestep_embedding, LVM/model.py:39-58): [cos | sin] of
// t * freqs in fp32, cast to bf16.  `freqs` (dim/2 fp32) comes from the host so that it is the
// same exp() the reference evaluates.
// ---------------------------------------------------------------------------------------------
__global__ void timestep_sinusoid_kernel(const float* __restrict__ t, const float* __restrict__ freqs,
                                         __nv_bfloat16* __restrict__ out, int n, int half) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n * half) return;
  const int r = idx / half, i = idx % half;
  const float ang = t[r] * freqs[i];
  out[(size_t)r * 2 * half + i] = __float2bfloat16_rn(cosf(ang));
  out[(size_t)r * 2 * half + half + i] = __float2bfloat16_rn(sinf(ang));
}

int timestep_sinusoid(const float* t, const float* freqs, void* out, int n, int dim, cudaStream_t s) {
  VGPT_CHECK_ARG(t && freqs && out && n >= 0 && dim > 0 && dim % 2 == 0,
                 "vgpt_timestep_sinusoid: bad arguments");
  if (n == 0) return 0;
  const int total = n * dim / 2;
  timestep_sinusoid_kernel<<<(total + 255) / 256, 256, 0, s>>>(t, freqs, (__nv_bfloat16*)out, n,
                                                              dim / 2);
  VGPT_CHECK_LAUNCH();
  return 0;
}

// ---------------------------------------------------------------------------------------------
// Small-batch linear: out[n,N] = post( pre(in[n,K]) @ W[N,K]^T + bias ), n <= 16 (TimestepEmbedder MLPs and adaLN
// modulation: LVM/model.py:32-36, 74-77; one row under the sampler's uniform timestep).  A weight-streaming GEMV: bound
// by the single pass over W (78.6 MB per Euler step for the five launches).  One warp per TWO output columns, every lane
// keeps eight 16-byte weight loads in flight (two columns x four k-chunks), inputs staged once per CTA in shared
// memory; R = rows rounded up to a power of two keeps the accumulators in registers.  (The first version -- one column
// per warp, one load in flight, 16 row slots whatever n -- reached 0.37 TB/s.)
// ---------------------------------------------------------------------------------------------
constexpr int kSmallMaxRows = 16;
constexpr int kSmallWarps = 8;
constexpr int kSmallCols = 2;          // output columns per warp

__device__ __forceinline__ float dot8(const uint4& w, const uint4& x, float acc) {
  acc = fmaf(bf16lo(w.x), bf16lo(x.x), acc); acc = fmaf(bf16hi(w.x), bf16hi(x.x), acc);
  acc = fmaf(bf16lo(w.y), bf16lo(x.y), acc); acc = fmaf(bf16hi(w.y), bf16hi(x.y), acc);
  acc = fmaf(bf16lo(w.z), bf16lo(x.z), acc); acc = fmaf(bf16hi(w.z), bf16hi(x.z), acc);
  acc = fmaf(bf16lo(w.w), bf16lo(x.w), acc); acc = fmaf(bf16hi(w.w), bf16hi(x.w), acc);
  return acc;
}

template <int R>
__global__ void __launch_bounds__(kSmallWarps * 32)
linear_small_kernel(const __nv_bfloat16* __restrict__ in, const __nv_bfloat16* __restrict__ W,
                    const __nv_bfloat16* __restrict__ bias, __nv_bfloat16* __restrict__ out, int n,
                    int N, int K, int pre_silu, int post_silu) {
  extern __shared__ __nv_bfloat16 sin_[];                 // [R][K] bf16 (after optional SiLU; rows >= n are zero)
  for (int i = threadIdx.x; i < R * K; i += blockDim.x) {
    float v = i < n * K ? __bfloat162float(in[i]) : 0.f;
    if (pre_silu) v = rbf(silu_f(v));
    sin_[i] = __float2bfloat16_rn(v);
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int col0 = (blockIdx.x * kSmallWarps + warp) * kSmallCols;
  if (col0 >= N) return;
  const bool two = col0 + 1 < N;
  float acc[R][kSmallCols];
#pragma unroll
  for (int r = 0; r < R; ++r) acc[r][0] = acc[r][1] = 0.f;
  const int kch = K >> 3;
  const uint4* w0 = reinterpret_cast<const uint4*>(W + (size_t)col0 * K);
  const uint4* w1 = reinterpret_cast<const uint4*>(W + (size_t)(two ? col0 + 1 : col0) * K);
  const uint4* xs = reinterpret_cast<const uint4*>(sin_);
  for (int c0 = lane; c0 < kch; c0 += 4 * 32) {
    uint4 wa[4], wb[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {                         // eight independent loads before the first use
      const int c = c0 + u * 32;
      wa[u] = c < kch ? __ldg(w0 + c) : make_uint4(0, 0, 0, 0);
      wb[u] = c < kch ? __ldg(w1 + c) : make_uint4(0, 0, 0, 0);
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int c = c0 + u * 32;
      if (c < kch) {
#pragma unroll
        for (int r = 0; r < R; ++r) {
          const uint4 xv = xs[(size_t)r * kch + c];
          acc[r][0] = dot8(wa[u], xv, acc[r][0]);
          acc[r][1] = dot8(wb[u], xv, acc[r][1]);
        }
      }
    }
  }
#pragma unroll
  for (int r = 0; r < R; ++r) {
#pragma unroll
    for (int j = 0; j < kSmallCols; ++j) {
      float v = warp_sum(acc[r][j]);
      if (lane == 0 && r < n && (j == 0 || two)) {
        v = rbf(v + (bias ? __bfloat162float(bias[col0 + j]) : 0.f));
        if (post_silu) v = rbf(silu_f(v));
        out[(size_t)r * N + col0 + j] = __float2bfloat16_rn(v);
      }
    }
  }
}

int linear_small(const void* in, const void* W, const void* bias, void* out, int n, int N, int K,
                 int pre_silu, int post_silu, cudaStream_t s) {
  VGPT_CHECK_ARG(in && W && out, "vgpt_linear_small: null pointer");
  VGPT_CHECK_ARG(n >= 0 && n <= kSmallMaxRows && N > 0 && K > 0 && K % 8 == 0,
                 "vgpt_linear_small: unsupported n=%d N=%d K=%d (n <= %d, K %% 8 == 0)", n, N, K,
                 kSmallMaxRows);
  VGPT_CHECK_ARG(((uintptr_t)in & 15) == 0 && ((uintptr_t)W & 15) == 0, "vgpt_linear_small: pointers must be 16-byte aligned");
  if (n == 0) return 0;
  const int R = n <= 1 ? 1 : n <= 2 ? 2 : n <= 4 ? 4 : n <= 8 ? 8 : 16;
  const size_t smem = (size_t)R * K * sizeof(__nv_bfloat16);
  const int grid = (N + kSmallWarps * kSmallCols - 1) / (kSmallWarps * kSmallCols);
#define VGPT_SMALL_CASE(R_)                                                                                              \
  if (R == R_) {                                                                                                         \
    if (smem > 48 * 1024)       /* per launch: the attribute is per device (cheap, capture-safe) */                      \
      VGPT_CHECK_CUDA(cudaFuncSetAttribute(linear_small_kernel<R_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    linear_small_kernel<R_><<<grid, kSmallWarps * 32, smem, s>>>((const __nv_bfloat16*)in, (const __nv_bfloat16*)W,     \
                                                                (const __nv_bfloat16*)bias, (__nv_bfloat16*)out, n, N, K, \
                                                                pre_silu, post_silu);                                     \
  }
  VGPT_SMALL_CASE(1)
  VGPT_SMALL_CASE(2)
  VGPT_SMALL_CASE(4)
  VGPT_SMALL_CASE(8)
  VGPT_SMALL_CASE(16)
#undef VGPT_SMALL_CASE
  VGPT_CHECK_LAUNCH();
  return 0;
}

// ---------------------------------------------------------------------------------------------
// CFG + flow-matching Euler update (LVM/scheduler.py:178-204, LVM/model.py:554-562) of ONE element pair.
//   x1 mode : v = (pred - z) / (1 - sigma)  for both branches, then CFG on v
//   v  mode : CFG on pred
//   CFG     : c = u + g * (c - u); both halves take c (the reference returns cond + cond)
//   Euler   : z += (sigma_next - sigma) * v
// Shared by cfg_euler_kernel and the fused final-layer kernel so that both round identically.
// ---------------------------------------------------------------------------------------------
struct StepScalars { float one_minus_sigma, dsigma, guidance; };

__device__ __forceinline__ void cfg_euler_element(__nv_bfloat16* __restrict__ z, float c, float u,
                                                  __nv_bfloat16* __restrict__ vel_out, int i, int half_numel,
                                                  int use_cfg, int x1_mode, const StepScalars& sc) {
  // torch divides a CUDA tensor by a CPU scalar as a * (1 / b) in fp32 (BinaryDivTrueKernel.cu);
  // the reference's `(pred - z) / (1.0 - sigma)` (scheduler.py:184) therefore rounds this way.
  const float inv = 1.0f / sc.one_minus_sigma;
  const float zc = __bfloat162float(z[i]);
  if (x1_mode) c = rbf(rbf(c - zc) * inv);
  float zu = 0.f;
  if (use_cfg) {
    zu = __bfloat162float(z[half_numel + i]);
    if (x1_mode) u = rbf(rbf(u - zu) * inv);
    c = rbf(u + rbf(sc.guidance * rbf(c - u)));
  }
  if (vel_out) vel_out[i] = __float2bfloat16_rn(c);
  const float step = rbf(sc.dsigma * c);
  z[i] = __float2bfloat16_rn(zc + step);
  if (use_cfg) z[half_numel + i] = __float2bfloat16_rn(zu + step);
}

// ---------------------------------------------------------------------------------------------
// Final layer (llm.norm + FinalLayer.forward + unpatchify, OmniGen/transformer.py:214, LVM/model.py:79-83,
// 255-265, 478-486) -- and, in the sampler's fused loop, the scheduler update behind it -- per image token
// of latent j:  [Phi3RMSNorm ->] LN(no affine, eps 1e-6) -> * (1 + scale_j) + shift_j -> Linear(h -> p*p*C)
// -> scatter feature (p,q,c) of patch (py,px) to pred[j][c][2py+p][2px+q]  [-> x1 -> v, CFG, Euler on z].
// One CTA per token (per cond / uncond token pair when the Euler update is fused).
// ---------------------------------------------------------------------------------------------
struct EulerFuse {          // z == nullptr: not fused
  __nv_bfloat16* z;
  __nv_bfloat16* vel;       // optional: applied velocity of the cond half
  const StepScalars* sc;    // device scalars (the step's 1 - sigma, d sigma, guidance): graph replayable
  int n_half;               // latents per CFG branch
  int use_cfg, x1_mode;
};

// the 16 output features of one row; valid in threads 0..15 (fp32, bias added, before the bf16 rounding)
__device__ __forceinline__ float final_row(const __nv_bfloat16* __restrict__ hidden, int row, int hs,
                                           const __nv_bfloat16* __restrict__ norm_w, float rms_eps,
                                           const __nv_bfloat16* __restrict__ mod_j, const __nv_bfloat16* __restrict__ w,
                                           const __nv_bfloat16* __restrict__ bias, float* red,
                                           float (*outs)[kRowThreads / 32]) {
  const int chunks = hs >> 3;
  const uint4* xr = reinterpret_cast<const uint4*>(hidden + (size_t)row * hs);
  float xv[kMaxChunksPerThread][8];
#pragma unroll
  for (int i = 0; i < kMaxChunksPerThread; ++i) {
    const int c = threadIdx.x + i * kRowThreads;
    if (c < chunks) unpack8(xr[c], xv[i]);
  }
  if (norm_w != nullptr) {             // llm.norm (Phi3RMSNorm): y = w * bf16(x * rstd), rounded to bf16
    float ss = 0.f;
#pragma unroll
    for (int i = 0; i < kMaxChunksPerThread; ++i) {
      const int c = threadIdx.x + i * kRowThreads;
      if (c < chunks) {
#pragma unroll
        for (int t = 0; t < 8; ++t) ss += xv[i][t] * xv[i][t];
      }
    }
    const float rstd = rsqrtf(block_sum(ss, red) / (float)hs + rms_eps);
#pragma unroll
    for (int i = 0; i < kMaxChunksPerThread; ++i) {
      const int c = threadIdx.x + i * kRowThreads;
      if (c < chunks) {
        float g[8];
        unpack8(reinterpret_cast<const uint4*>(norm_w)[c], g);
#pragma unroll
        for (int t = 0; t < 8; ++t) xv[i][t] = rbf(g[t] * rbf(xv[i][t] * rstd));
      }
    }
  }
  float s1 = 0.f;
#pragma unroll
  for (int i = 0; i < kMaxChunksPerThread; ++i) {
    const int c = threadIdx.x + i * kRowThreads;
    if (c < chunks) {
#pragma unroll
      for (int t = 0; t < 8; ++t) s1 += xv[i][t];
    }
  }
  const float mean = block_sum(s1, red) / (float)hs;
  float s2 = 0.f;
#pragma unroll
  for (int i = 0; i < kMaxChunksPerThread; ++i) {
    const int c = threadIdx.x + i * kRowThreads;
    if (c < chunks) {
#pragma unroll
      for (int t = 0; t < 8; ++t) { const float d = xv[i][t] - mean; s2 += d * d; }
    }
  }
  const float rstd = rsqrtf(block_sum(s2, red) / (float)hs + 1e-6f);
  const uint4* shift = reinterpret_cast<const uint4*>(mod_j);
  const uint4* scale = reinterpret_cast<const uint4*>(mod_j + hs);
  float acc[16];
#pragma unroll
  for (int f = 0; f < 16; ++f) acc[f] = 0.f;
#pragma unroll
  for (int i = 0; i < kMaxChunksPerThread; ++i) {
    const int c = threadIdx.x + i * kRowThreads;
    if (c < chunks) {
      float sh[8], sc[8];
      unpack8(shift[c], sh);
      unpack8(scale[c], sc);
#pragma unroll
      for (int t = 0; t < 8; ++t) {
        float y = rbf((xv[i][t] - mean) * rstd);          // layer_norm output (bf16)
        y = rbf(y * rbf(1.0f + sc[t]));                   // x * (1 + scale)
        xv[i][t] = rbf(y + sh[t]);                        // + shift
      }
#pragma unroll
      for (int f = 0; f < 16; ++f) {
        float wv[8];
        unpack8(reinterpret_cast<const uint4*>(w + (size_t)f * hs)[c], wv);
#pragma unroll
        for (int t = 0; t < 8; ++t) acc[f] = fmaf(wv[t], xv[i][t], acc[f]);
      }
    }
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  __syncthreads();                                        // outs[] may still be read from the previous row
#pragma unroll
  for (int f = 0; f < 16; ++f) {
    const float v = warp_sum(acc[f]);
    if (lane == 0) outs[f][warp] = v;
  }
  __syncthreads();
  float v = 0.f;
  if (threadIdx.x < 16) {
#pragma unroll
    for (int k = 0; k < kRowThreads / 32; ++k) v += outs[threadIdx.x][k];
    v += __bfloat162float(bias[threadIdx.x]);
  }
  return v;
}

__global__ void __launch_bounds__(kRowThreads)
final_layer_kernel(const __nv_bfloat16* __restrict__ hidden, int hs, const __nv_bfloat16* __restrict__ norm_w,
                   float rms_eps, const int32_t* __restrict__ lat_row0, const int32_t* __restrict__ row_kind,
                   const int32_t* __restrict__ row_a, const int32_t* __restrict__ row_b,
                   const __nv_bfloat16* __restrict__ mod, const __nv_bfloat16* __restrict__ w,
                   const __nv_bfloat16* __restrict__ bias, PeerPtrs preds, int n_preds,
                   int tokens_per_lat, int C, int lat_h, int lat_w, EulerFuse eu) {
  __shared__ float red[32];
  __shared__ float outs[16][kRowThreads / 32];
  int j, tkn, row;
  if (lat_row0) {                    // latent-driven: CTA = (latent, token), rows contiguous per latent
    j = blockIdx.x / tokens_per_lat; tkn = blockIdx.x % tokens_per_lat;
    row = lat_row0[j] + tkn;
  } else {                           // row-driven (any row partition): CTA = row, image-token rows only
    row = blockIdx.x;
    if (row_kind[row] != 2 /* VGPT_ROW_NOISY_PATCH */) return;
    j = row_a[row]; tkn = row_b[row];
  }
  const float v = final_row(hidden, row, hs, norm_w, rms_eps, mod + (size_t)j * 2 * hs, w, bias, red, outs);
  float vu = 0.f;
  if (eu.z != nullptr && eu.use_cfg) {                    // the unconditional twin of this token (latent j + n_half)
    const int ju = j + eu.n_half;
    vu = final_row(hidden, lat_row0[ju] + tkn, hs, norm_w, rms_eps, mod + (size_t)ju * 2 * hs, w, bias, red, outs);
  }
  if (threadIdx.x < 16) {
    const int f = threadIdx.x;
    const int pw = lat_w >> 1;
    const int py = tkn / pw, px = tkn % pw;
    const int p = f / (2 * C), q = (f / C) & 1, c = f % C;     // feature order (p, q, c)
    const size_t numel_lat = (size_t)C * lat_h * lat_w;
    const size_t off = ((size_t)c * lat_h + 2 * py + p) * lat_w + 2 * px + q;
    const __nv_bfloat16 r = __float2bfloat16_rn(v);
#pragma unroll
    for (int g = 0; g < kMaxPeers; ++g)
      if (g < n_preds) static_cast<__nv_bfloat16*>(preds.p[g])[(size_t)j * numel_lat + off] = r;
    if (eu.z != nullptr) {
      const __nv_bfloat16 ru = __float2bfloat16_rn(vu);
      if (eu.use_cfg) static_cast<__nv_bfloat16*>(preds.p[0])[(size_t)(j + eu.n_half) * numel_lat + off] = ru;
      const StepScalars sc = *eu.sc;
      cfg_euler_element(eu.z, __bfloat162float(r), __bfloat162float(ru), eu.vel, (int)((size_t)j * numel_lat + off),
                        (int)(eu.n_half * numel_lat), eu.use_cfg, eu.x1_mode, sc);
    }
  }
}

int final_layer(const void* hidden, int hs, const void* norm_w, float rms_eps, const int32_t* lat_row0, const void* mod,
                const void* w, const void* bias, void* pred, int n_lat, int C, int lat_h, int lat_w, void* z_euler,
                void* vel_out, const float* scalars_dev, int use_cfg, int x1_mode, cudaStream_t s) {
  VGPT_CHECK_ARG(hidden && lat_row0 && mod && w && bias && pred, "vgpt_final_layer: null pointer");
  VGPT_CHECK_ARG(hs % 8 == 0 && hs <= kRowThreads * kMaxChunksPerThread * 8 && C == 4 &&
                     lat_h % 2 == 0 && lat_w % 2 == 0,
                 "vgpt_final_layer: unsupported hs=%d C=%d latent %dx%d", hs, C, lat_h, lat_w);
  VGPT_CHECK_ARG(!z_euler || (scalars_dev && (!use_cfg || n_lat % 2 == 0)),
                 "vgpt_final_layer: the fused scheduler update needs device scalars and, with CFG, an even number of latents");
  if (n_lat <= 0) return 0;
  const int tokens = (lat_h / 2) * (lat_w / 2);
  PeerPtrs pp = {};
  pp.p[0] = pred;
  EulerFuse eu = {};
  int ctas = n_lat * tokens;
  if (z_euler) {
    eu.z = static_cast<__nv_bfloat16*>(z_euler);
    eu.vel = static_cast<__nv_bfloat16*>(vel_out);
    eu.sc = reinterpret_cast<const StepScalars*>(scalars_dev);
    eu.use_cfg = use_cfg;
    eu.x1_mode = x1_mode;
    eu.n_half = use_cfg ? n_lat / 2 : n_lat;
    ctas = eu.n_half * tokens;            // one CTA per cond / uncond token pair
  }
  final_layer_kernel<<<ctas, kRowThreads, 0, s>>>(
      (const __nv_bfloat16*)hidden, hs, (const __nv_bfloat16*)norm_w, rms_eps, lat_row0, nullptr, nullptr, nullptr,
      (const __nv_bfloat16*)mod, (const __nv_bfloat16*)w, (const __nv_bfloat16*)bias, pp, 1, tokens, C, lat_h, lat_w, eu);
  VGPT_CHECK_LAUNCH();
  return 0;
}

// Row-driven variant for row-sharded (sequence-parallel) plans: `rows` local rows described by the
// plan's (kind, a = latent, b = token) arrays; the prediction is stored into every rank's pred
// buffer (preds[0..n_preds), NVLink stores), so no gather follows.
int final_layer_rows(const void* hidden, int rows, int hs, const void* norm_w, float rms_eps, const int32_t* kind,
                     const int32_t* a, const int32_t* b, const void* mod, const void* w, const void* bias,
                     void* const* preds, int n_preds, int C, int lat_h, int lat_w, cudaStream_t s) {
  VGPT_CHECK_ARG(hidden && kind && a && b && mod && w && bias && preds, "vgpt_final_layer_rows: null pointer");
  VGPT_CHECK_ARG(n_preds >= 1 && n_preds <= kMaxPeers, "vgpt_final_layer_rows: %d destinations (1..%d)", n_preds, kMaxPeers);
  VGPT_CHECK_ARG(hs % 8 == 0 && hs <= kRowThreads * kMaxChunksPerThread * 8 && C == 4 &&
                     lat_h % 2 == 0 && lat_w % 2 == 0,
                 "vgpt_final_layer_rows: unsupported hs=%d C=%d latent %dx%d", hs, C, lat_h, lat_w);
  PeerPtrs pp;
  for (int i = 0; i < kMaxPeers; ++i) {
    pp.p[i] = i < n_preds ? preds[i] : nullptr;
    VGPT_CHECK_ARG(i >= n_preds || pp.p[i], "vgpt_final_layer_rows: null destination %d", i);
  }
  if (rows <= 0) return 0;
  final_layer_kernel<<<rows, kRowThreads, 0, s>>>(
      (const __nv_bfloat16*)hidden, hs, (const __nv_bfloat16*)norm_w, rms_eps, nullptr, kind, a, b,
      (const __nv_bfloat16*)mod, (const __nv_bfloat16*)w, (const __nv_bfloat16*)bias, pp, n_preds,
      (lat_h / 2) * (lat_w / 2), C, lat_h, lat_w, EulerFuse{});
  VGPT_CHECK_LAUNCH();
  return 0;
}

// ---------------------------------------------------------------------------------------------
// Stand-alone scheduler update (any callback through the S2 seam, CFG-branch pairs, sequence-parallel groups).
// z, pred: [n_branches * n_cond * numel] bf16 laid out [cond latents..., uncond latents...].
// ---------------------------------------------------------------------------------------------
__global__ void cfg_euler_kernel(__nv_bfloat16* __restrict__ z, const __nv_bfloat16* __restrict__ pred,
                                 __nv_bfloat16* __restrict__ vel_out, int half_numel, int use_cfg,
                                 int x1_mode, const StepScalars* __restrict__ sc_dev, StepScalars sc_host) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= half_numel) return;
  const StepScalars sc = sc_dev ? *sc_dev : sc_host;
  cfg_euler_element(z, __bfloat162float(pred[i]), use_cfg ? __bfloat162float(pred[half_numel + i]) : 0.f, vel_out, i,
                    half_numel, use_cfg, x1_mode, sc);
}

int cfg_euler(void* z, const void* pred, void* vel_out, int half_numel, int use_cfg, int x1_mode,
              float one_minus_sigma, float dsigma, float guidance, const float* scalars_dev,
              cudaStream_t s) {
  VGPT_CHECK_ARG(z && pred && half_numel >= 0, "vgpt_cfg_euler: bad arguments");
  if (half_numel == 0) return 0;
  StepScalars sc{one_minus_sigma, dsigma, guidance};
  cfg_euler_kernel<<<(half_numel + 255) / 256, 256, 0, s>>>(
      (__nv_bfloat16*)z, (const __nv_bfloat16*)pred, (__nv_bfloat16*)vel_out, half_numel, use_cfg,
      x1_mode, reinterpret_cast<const StepScalars*>(scalars_dev), sc);
  VGPT_CHECK_LAUNCH();
  return 0;
}

// v-mode CFG inside the model (LVM/model.py:554-562): cond = uncond + g * (cond - uncond), and the
// model returns cond + cond, so both halves of `pred` receive the combined value.
__global__ void cfg_combine_kernel(__nv_bfloat16* __restrict__ pred, int half_numel, float guidance) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= half_numel) return;
  const float c = __bfloat162float(pred[i]), u = __bfloat162float(pred[half_numel + i]);
  const __nv_bfloat16 r = __float2bfloat16_rn(u + rbf(guidance * rbf(c - u)));
  pred[i] = r;
  pred[half_numel + i] = r;
}

int cfg_combine(void* pred, int half_numel, float guidance, cudaStream_t s) {
  VGPT_CHECK_ARG(pred && half_numel >= 0, "vgpt_cfg_combine: bad arguments");
  if (half_numel == 0) return 0;
  cfg_combine_kernel<<<(half_numel + 255) / 256, 256, 0, s>>>((__nv_bfloat16*)pred, half_numel, guidance);
  VGPT_CHECK_LAUNCH();
  return 0;
}

// ---------------------------------------------------------------------------------------------
// SwiGLU weight repack: gate_up_proj [2I, K] = [gate(I) | up(I)] rows  ->  per 32 packed rows
// [gate x16 | up x16] of 16 consecutive outputs, so one 32-column accumulator chunk holds matching pairs (and, for
// the transposed tail accumulator of gemm_pair_tcgen05.cu, one 32-lane TMEM quadrant does).
// ---------------------------------------------------------------------------------------------
__global__ void pack_gate_up_kernel(const uint4* __restrict__ w, uint4* __restrict__ out, int I,
                                    int kchunks) {
  const int prow = blockIdx.x;                       // packed row
  const int blk = prow >> 5, r = prow & 31;
  const int src = (r < 16) ? (blk * 16 + r) : (I + blk * 16 + (r - 16));
  for (int c = threadIdx.x; c < kchunks; c += blockDim.x)
    out[(size_t)prow * kchunks + c] = w[(size_t)src * kchunks + c];
}

int pack_gate_up(const void* w, void* packed, int I, int K, cudaStream_t s) {
  VGPT_CHECK_ARG(w && packed && w != packed, "vgpt_pack_gate_up: bad pointers");
  VGPT_CHECK_ARG(I > 0 && I % 32 == 0 && K % 8 == 0, "vgpt_pack_gate_up: I=%d K=%d unsupported", I, K);
  pack_gate_up_kernel<<<2 * I, 128, 0, s>>>((const uint4*)w, (uint4*)packed, I, K / 8);
  VGPT_CHECK_LAUNCH();
  return 0;
}

// ---------------------------------------------------------------------------------------------
// Dense mask from token codes: out[q][k] = (q_code[q] >= k_code[k]).  The attention kernel
// evaluates exactly this predicate in registers; materialising it is only for parity tests and
// for callers that want the reference's [L,L] tensor (LVM/processor.py:682-731).
// ---------------------------------------------------------------------------------------------
__global__ void mask_from_codes_kernel(const int32_t* __restrict__ qc, const int32_t* __restrict__ kc,
                                       uint8_t* __restrict__ out, int Lq, int Lk) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  const int q = blockIdx.y;
  if (k < Lk) out[(size_t)q * Lk + k] = qc[q] >= kc[k] ? 1 : 0;
}

int mask_from_codes(const int32_t* qc, const int32_t* kc, void* out, int Lq, int Lk, cudaStream_t s) {
  VGPT_CHECK_ARG(qc && kc && out && Lq >= 0 && Lk >= 0, "vgpt_mask_from_codes: bad arguments");
  if (Lq == 0 || Lk == 0) return 0;
  dim3 grid((Lk + 255) / 256, Lq);
  mask_from_codes_kernel<<<grid, 256, 0, s>>>(qc, kc, (uint8_t*)out, Lq, Lk);
  VGPT_CHECK_LAUNCH();
  return 0;
}

}  // namespace vgpt
