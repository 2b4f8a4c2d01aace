// Descriptor probe: runs a handful of tcgen05.mma instructions on caller-supplied raw shared
// memory images and descriptor fields, and returns the fp32 accumulator tile.  It exists so the
// shared-memory layouts / descriptor encodings the production kernels rely on (K-major and
// MN-major operands, 32/64/128-byte swizzles, K-advance strides) are pinned by a test on the GPU
// (tests/test_umma_layouts.py) instead of being assumed.  Built into libvgpt_b200_probe.so (tests and
// tools/umma_rate.py only), NOT into the product library.
#include "../common.cuh"
#include "../vgpt_internal.h"

namespace vgpt {

int umma_probe_ts(const void* a_words, int a_cols, const void* b_img, int b_bytes, uint64_t b_desc_base,
                  uint32_t idesc, int k_steps, uint32_t b_step_bytes, float* d_out, int n_cols, cudaStream_t s);
int umma_probe(const void* a_img, int a_bytes, const void* b_img, int b_bytes, uint64_t a_desc_base,
               uint64_t b_desc_base, uint32_t idesc, int k_steps, uint32_t a_step_bytes,
               uint32_t b_step_bytes, float* d_out, int n_cols, cudaStream_t s);
int umma_rate(int mode, int N, int iters, int n_acc, int commit_every, int ctas, float* out, cudaStream_t s);

__global__ void __launch_bounds__(128, 1)
umma_probe_kernel(const uint4* __restrict__ a_img, int a_bytes, const uint4* __restrict__ b_img,
                  int b_bytes, uint64_t a_desc_base, uint64_t b_desc_base, uint32_t idesc,
                  int k_steps, uint32_t a_step_bytes, uint32_t b_step_bytes,
                  float* __restrict__ d_out, int n_cols) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gen = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t a_off = 0, b_off = (uint32_t)((a_bytes + 1023) & ~1023);
  for (int i = threadIdx.x; i < a_bytes / 16; i += blockDim.x)
    reinterpret_cast<uint4*>(gen + a_off)[i] = a_img[i];
  for (int i = threadIdx.x; i < b_bytes / 16; i += blockDim.x)
    reinterpret_cast<uint4*>(gen + b_off)[i] = b_img[i];
  fence_proxy_async_smem();          // generic-proxy smem writes -> visible to the tensor core
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    mbar_init(smem_u32(&bar), 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc<256>(smem_u32(&tmem_slot));
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  if (threadIdx.x == 0) {
    for (int k = 0; k < k_steps; ++k) {
      const uint64_t da = a_desc_base + (uint64_t)(((base + a_off + k * a_step_bytes) >> 4) & 0x3fffu);
      const uint64_t db = b_desc_base + (uint64_t)(((base + b_off + k * b_step_bytes) >> 4) & 0x3fffu);
      umma_f16_ss(tmem, da, db, idesc, k > 0 ? 1u : 0u);
    }
    umma_commit(smem_u32(&bar));
  }
  mbar_wait(smem_u32(&bar), 0);
  tc_fence_after();
  const int row = warp * 32 + lane;
  for (int c = 0; c < n_cols; c += 32) {
    uint32_t v[32];
    tmem_ld_32x32b_x32(tmem + ((uint32_t)(warp * 32) << 16) + c, v);
    tmem_ld_wait();
    for (int j = 0; j < 32 && c + j < n_cols; ++j) d_out[(size_t)row * n_cols + c + j] = __uint_as_float(v[j]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc<256>(tmem);
  }
}

// Same probe with the A operand in tensor memory: a_words[128][k_steps * 8] are the packed bf16x2
// words of each row (thread r stores row r with tcgen05.st), B from a raw smem image.
__global__ void __launch_bounds__(128, 1)
umma_probe_ts_kernel(const uint32_t* __restrict__ a_words, int a_cols, const uint4* __restrict__ b_img,
                     int b_bytes, uint64_t b_desc_base, uint32_t idesc, int k_steps, uint32_t b_step_bytes,
                     float* __restrict__ d_out, int n_cols) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gen = smem_raw + (base - smem_u32(smem_raw));
  for (int i = threadIdx.x; i < b_bytes / 16; i += blockDim.x) reinterpret_cast<uint4*>(gen)[i] = b_img[i];
  fence_proxy_async_smem();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    mbar_init(smem_u32(&bar), 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc<512>(smem_u32(&tmem_slot));
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  const uint32_t tmem_a = tmem + 256;
  const int row = warp * 32 + lane;
  for (int c = 0; c < a_cols; c += 32) {
    uint32_t v[32];
    for (int j = 0; j < 32; ++j) v[j] = (c + j < a_cols) ? a_words[(size_t)row * a_cols + c + j] : 0u;
    tmem_st_32x32b_x32(tmem_a + ((uint32_t)(warp * 32) << 16) + c, v);
  }
  tmem_st_wait();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (threadIdx.x == 0) {
    for (int k = 0; k < k_steps; ++k) {
      const uint64_t db = b_desc_base + (uint64_t)(((base + k * b_step_bytes) >> 4) & 0x3fffu);
      umma_f16_ts(tmem, tmem_a + k * 8, db, idesc, k > 0 ? 1u : 0u);
    }
    umma_commit(smem_u32(&bar));
  }
  mbar_wait(smem_u32(&bar), 0);
  tc_fence_after();
  for (int c = 0; c < n_cols; c += 32) {
    uint32_t v[32];
    tmem_ld_32x32b_x32(tmem + ((uint32_t)(warp * 32) << 16) + c, v);
    tmem_ld_wait();
    for (int j = 0; j < 32 && c + j < n_cols; ++j) d_out[(size_t)row * n_cols + c + j] = __uint_as_float(v[j]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc<512>(tmem);
  }
}

int umma_probe_ts(const void* a_words, int a_cols, const void* b_img, int b_bytes, uint64_t b_desc_base,
                  uint32_t idesc, int k_steps, uint32_t b_step_bytes, float* d_out, int n_cols, cudaStream_t s) {
  VGPT_CHECK_ARG(a_words && b_img && d_out, "vgpt_debug_umma_probe_ts: null pointer");
  VGPT_CHECK_ARG(a_cols > 0 && a_cols <= 128 && b_bytes > 0 && b_bytes % 16 == 0 && b_bytes <= 96 * 1024 &&
                     n_cols > 0 && n_cols <= 256 && k_steps > 0 && k_steps * 8 <= a_cols,
                 "vgpt_debug_umma_probe_ts: bad sizes");
  const int smem = ((b_bytes + 1023) & ~1023) + 1024;
  VGPT_CHECK_CUDA(cudaFuncSetAttribute(umma_probe_ts_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  umma_probe_ts_kernel<<<1, 128, smem, s>>>((const uint32_t*)a_words, a_cols, (const uint4*)b_img, b_bytes,
                                            b_desc_base, idesc, k_steps, b_step_bytes, d_out, n_cols);
  VGPT_CHECK_LAUNCH();
  return 0;
}

int umma_probe(const void* a_img, int a_bytes, const void* b_img, int b_bytes, uint64_t a_desc_base,
               uint64_t b_desc_base, uint32_t idesc, int k_steps, uint32_t a_step_bytes,
               uint32_t b_step_bytes, float* d_out, int n_cols, cudaStream_t s) {
  VGPT_CHECK_ARG(a_img && b_img && d_out, "vgpt_debug_umma_probe: null pointer");
  VGPT_CHECK_ARG(a_bytes > 0 && b_bytes > 0 && a_bytes % 16 == 0 && b_bytes % 16 == 0 &&
                     a_bytes <= 96 * 1024 && b_bytes <= 96 * 1024,
                 "vgpt_debug_umma_probe: image sizes must be multiples of 16 and <= 96 KiB");
  VGPT_CHECK_ARG(n_cols > 0 && n_cols <= 256 && k_steps > 0, "vgpt_debug_umma_probe: bad n_cols / k_steps");
  const int smem = ((a_bytes + 1023) & ~1023) + ((b_bytes + 1023) & ~1023) + 1024;
  VGPT_CHECK_CUDA(cudaFuncSetAttribute(umma_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  umma_probe_kernel<<<1, 128, smem, s>>>((const uint4*)a_img, a_bytes, (const uint4*)b_img, b_bytes,
                                         a_desc_base, b_desc_base, idesc, k_steps, a_step_bytes,
                                         b_step_bytes, d_out, n_cols);
  VGPT_CHECK_LAUNCH();
  return 0;
}


// ---------------------------------------------------------------------------------------------
// Rate probe: how many SM cycles does one tcgen05.mma (M = 128, cta_group::1) of a given shape and
// operand form really take when issued back to back?  One CTA per SM, operands resident (zeros),
// no loads.  mode: 0 = SS, K-major, SW128; 1 = SS, K-major, SW64; 2 = TS (A in TMEM), B MN-major
// SW128; 3 = TS, B MN-major SW64.  n_acc = 1: every MMA accumulates into the same tile (dependent
// chain); 2: two tiles alternate.  out[cta] = cycles per MMA (clock64 around issue + completion).
// Not on the hot path (tools/umma_rate.py; numbers in DESIGN.md).
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128, 1)
umma_rate_kernel(int mode, int N, int iters, int n_acc, int commit_every, float* __restrict__ out) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ __align__(8) uint64_t bar2;
  __shared__ uint32_t tmem_slot;
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gen = smem_raw + (base - smem_u32(smem_raw));
  for (int i = threadIdx.x; i < (96 * 1024) / 16; i += blockDim.x) reinterpret_cast<uint4*>(gen)[i] = make_uint4(0, 0, 0, 0);
  fence_proxy_async_smem();
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); mbar_init(smem_u32(&bar2), 1 << 20); fence_barrier_init(); }
  if (warp == 0) tmem_alloc<512>(smem_u32(&tmem_slot));
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  long long t0 = 0, t1 = 0;
  if (warp == 0) {
    const uint32_t a_addr = base, b_addr = base + 32 * 1024;
    const bool ts = mode >= 2;
    const bool sw64 = (mode & 1) != 0;
    const uint32_t row_bytes = sw64 ? 64 : 128, layout = sw64 ? kLayoutSW64 : kLayoutSW128;
    const uint32_t idesc = make_idesc_bf16(128, N, 0, ts ? 1 : 0);
    const int ksteps_per_chunk = row_bytes / 32;
    t0 = clock64();
    for (int it = 0; it < iters; it += 8) {
      if (elect_one_sync()) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const int k = u % ksteps_per_chunk;
          const uint32_t d = tmem + ((n_acc == 2 && (u & 1)) ? 128u : 0u);
          if (!ts) {
            const uint64_t da = make_smem_desc(a_addr + k * 32, 16, 8 * row_bytes, layout);
            const uint64_t db = make_smem_desc(b_addr + k * 32, 16, 8 * row_bytes, layout);
            umma_f16_ss(d, da, db, idesc, 1u);
          } else {
            const uint64_t db = make_smem_desc(b_addr + (u & 3) * 16 * row_bytes, 128 * row_bytes, 8 * row_bytes, layout);
            umma_f16_ts(d, tmem + 256 + (u & 7) * 8, db, idesc, 1u);
          }
          if (commit_every > 0 && (u + 1) % commit_every == 0) umma_commit(smem_u32(&bar2));   // nobody waits on it
        }
      }
      __syncwarp();
    }
    if (elect_one_sync()) umma_commit(smem_u32(&bar));
    __syncwarp();
    mbar_wait(smem_u32(&bar), 0);
    t1 = clock64();
    if (threadIdx.x == 0) out[blockIdx.x] = (float)(t1 - t0) / (float)iters;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc<512>(tmem);
  }
}

// Same probe for CTA pairs (tcgen05.mma.cta_group::2, M = 256 over two SMs, SS form, K-major
// SW128): the production GEMM's instruction.  out[cluster] = cycles per MMA seen by the leader.
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1)
umma_rate_pair_kernel(int N, int iters, float* __restrict__ out) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gen = smem_raw + (base - smem_u32(smem_raw));
  for (int i = threadIdx.x; i < (64 * 1024) / 16; i += blockDim.x) reinterpret_cast<uint4*>(gen)[i] = make_uint4(0, 0, 0, 0);
  fence_proxy_async_smem();
  const int warp = threadIdx.x >> 5;
  const bool leader = cluster_ctarank() == 0;
  if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); fence_barrier_init(); }
  if (warp == 0) tmem_alloc_pair<512>(smem_u32(&tmem_slot));
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  if (warp == 0) {
    long long t0 = clock64();
    if (leader) {
      const uint32_t idesc = make_idesc_bf16(256, N);
      for (int it = 0; it < iters; it += 8) {
        if (elect_one_sync()) {
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            const uint64_t da = make_smem_desc(base + (u & 3) * 32, 16, 1024, kLayoutSW128);
            const uint64_t db = make_smem_desc(base + 32 * 1024 + (u & 3) * 32, 16, 1024, kLayoutSW128);
            umma_f16_ss_pair(tmem, da, db, idesc, 1u);
          }
        }
        __syncwarp();
      }
      if (elect_one_sync()) umma_commit_pair(smem_u32(&bar));
      __syncwarp();
    }
    mbar_wait(smem_u32(&bar), 0);
    const long long t1 = clock64();
    if (leader && threadIdx.x == 0) out[blockIdx.x >> 1] = (float)(t1 - t0) / (float)iters;
  }
  tc_fence_before();
  cluster_sync_all();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc_pair<512>(tmem);
  }
}

int umma_rate(int mode, int N, int iters, int n_acc, int commit_every, int ctas, float* out, cudaStream_t s) {
  VGPT_CHECK_ARG(out && mode >= 0 && mode <= 4 && N >= 16 && N <= 256 && N % 16 == 0 && iters >= 8 && iters % 8 == 0 &&
                     (n_acc == 1 || (n_acc == 2 && N <= 128)) && ctas >= 1 && iters <= (1 << 19) &&
                     (commit_every == 0 || 8 % commit_every == 0),
                 "vgpt_debug_umma_rate: bad arguments");
  const int smem = 96 * 1024 + 1024;
  VGPT_CHECK_CUDA(cudaFuncSetAttribute(umma_rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  if (mode == 4) {                       // CTA pairs: `ctas` = number of clusters
    const int smem2 = 64 * 1024 + 1024;
    VGPT_CHECK_CUDA(cudaFuncSetAttribute(umma_rate_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem2));
    umma_rate_pair_kernel<<<2 * ctas, 128, smem2, s>>>(N, iters, out);
    VGPT_CHECK_LAUNCH();
    return 0;
  }
  umma_rate_kernel<<<ctas, 128, smem, s>>>(mode, N, iters, n_acc, commit_every, out);
  VGPT_CHECK_LAUNCH();
  return 0;
}

}  // namespace vgpt

// ---------------------------------------------------------------------------------------------
// C ABI of the probe library (include/vgpt_b200_probe.h)
// ---------------------------------------------------------------------------------------------
extern "C" {
const char* vgpt_probe_last_error(void) { return vgpt::last_error(); }
int vgpt_debug_umma_probe(const void* a_img, int a_bytes, const void* b_img, int b_bytes, uint64_t a_desc_base,
                          uint64_t b_desc_base, uint32_t idesc, int k_steps, uint32_t a_step_bytes,
                          uint32_t b_step_bytes, float* d_out, int n_cols, void* stream) {
  return vgpt::umma_probe(a_img, a_bytes, b_img, b_bytes, a_desc_base, b_desc_base, idesc, k_steps, a_step_bytes,
                          b_step_bytes, d_out, n_cols, static_cast<cudaStream_t>(stream));
}
int vgpt_debug_umma_probe_ts(const void* a_words, int a_cols, const void* b_img, int b_bytes, uint64_t b_desc_base,
                             uint32_t idesc, int k_steps, uint32_t b_step_bytes, float* d_out, int n_cols, void* stream) {
  return vgpt::umma_probe_ts(a_words, a_cols, b_img, b_bytes, b_desc_base, idesc, k_steps, b_step_bytes, d_out, n_cols,
                             static_cast<cudaStream_t>(stream));
}
int vgpt_debug_umma_rate(int mode, int N, int iters, int n_acc, int commit_every, int ctas, float* out, void* stream) {
  return vgpt::umma_rate(mode, N, iters, n_acc, commit_every, ctas, out, static_cast<cudaStream_t>(stream));
}
}  // extern "C"
