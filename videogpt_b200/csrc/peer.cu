// Peer memory over NVLink 5 / NVSwitch: the plumbing of the sequence-parallel path.
//
// One process per GPU.  Every rank allocates its KV pools / prediction buffer / flag words with
// vgpt_peer_alloc (plain cudaMalloc, so the allocation can be exported as a CUDA IPC handle), the
// handles are exchanged on the host (torch.distributed, once per plan) and imported with
// vgpt_peer_import, after which a kernel running on GPU a can store straight into GPU b's HBM.
// The data-path kernels that PRODUCE something every rank needs (post-RoPE K / V rows, the
// final-layer prediction) write it to all peers themselves -- the all-gather is fused into the
// producing kernel, there is no staging buffer and no NCCL call on the data path -- and
// peer_barrier_kernel is the only synchronisation: a flag exchange through the same peer
// mappings (release store at system scope, acquire spin), one per layer.
//
// Replaces the reference's DeepSpeed-Ulysses all-to-alls (LVM/transform/sdpa_transform.py:126-156,
// four collectives per layer) and its hidden-state all-gather (LVM/model.py:466-474).
#include "common.cuh"
#include "vgpt_internal.h"

#include <cstring>

namespace vgpt {

int peer_alloc(void** out, uint64_t bytes) {
  VGPT_CHECK_ARG(out && bytes > 0, "vgpt_peer_alloc: bad arguments");
  void* p = nullptr;
  VGPT_CHECK_CUDA(cudaMalloc(&p, bytes));
  VGPT_CHECK_CUDA(cudaMemset(p, 0, bytes));
  VGPT_CHECK_CUDA(cudaDeviceSynchronize());
  *out = p;
  return 0;
}

int peer_free(void* p) {
  if (p) VGPT_CHECK_CUDA(cudaFree(p));
  return 0;
}

int peer_export(void* p, void* handle64) {
  VGPT_CHECK_ARG(p && handle64, "vgpt_peer_export: null pointer");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle is 64 bytes");
  cudaIpcMemHandle_t h;
  VGPT_CHECK_CUDA(cudaIpcGetMemHandle(&h, p));
  memcpy(handle64, &h, 64);
  return 0;
}

int peer_import(const void* handle64, void** out) {
  VGPT_CHECK_ARG(handle64 && out, "vgpt_peer_import: null pointer");
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, 64);
  void* p = nullptr;
  VGPT_CHECK_CUDA(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
  *out = p;
  return 0;
}

int peer_close(void* p) {
  if (p) VGPT_CHECK_CUDA(cudaIpcCloseMemHandle(p));
  return 0;
}

// ---------------------------------------------------------------------------------------------
// Barrier across the ranks of a group.  flags.p[j] = rank j's flag words (uint32[n], peer mapped;
// flags.p[rank] is local).  state (local uint32[4]): [0] = epoch counter, [1] = sticky time-out flag,
// [2..3] = uint64 nanoseconds (globaltimer) spent inside barrier kernels so far -- waiting for the
// slowest peer plus the NVLink flag round trip; bench.py reports it per Euler step.
// Epochs advance by one per call on every rank, so the kernel is replayable from a CUDA graph.
// All writes of earlier kernels in this stream (including stores into peer memory) are ordered
// before the flag by the kernel boundary + the system-scope fence / release store.
// ---------------------------------------------------------------------------------------------
__global__ void peer_barrier_kernel(PeerPtrs flags, int n, int rank, uint32_t* __restrict__ state,
                                    long long timeout_cycles) {
  __shared__ uint32_t s_epoch;
  unsigned long long t_in = 0;
  if (threadIdx.x == 0) {
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_in));
    s_epoch = state[0] + 1;
    state[0] = s_epoch;
  }
  __syncthreads();
  const uint32_t epoch = s_epoch;
  const int t = threadIdx.x;
  if (t < n) {
    __threadfence_system();
    uint32_t* remote = static_cast<uint32_t*>(flags.p[t]) + rank;
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(remote), "r"(epoch) : "memory");
    const uint32_t* mine = static_cast<const uint32_t*>(flags.p[rank]) + t;
    const long long t0 = clock64();
    const bool dead = *reinterpret_cast<volatile uint32_t*>(state + 1) != 0;   // fail fast after a time-out
    for (; !dead;) {
      uint32_t v;
      asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(mine) : "memory");
      if ((int32_t)(v - epoch) >= 0) break;
      if (clock64() - t0 > timeout_cycles) {          // a peer died: do not hang the GPU
        state[1] = 1;
        break;
      }
    }
    __threadfence_system();
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned long long t_out;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_out));
    *reinterpret_cast<unsigned long long*>(state + 2) += t_out - t_in;
  }
}

int peer_barrier(void* const* flag_ptrs, int n, int rank, uint32_t* state, cudaStream_t s) {
  VGPT_CHECK_ARG(flag_ptrs && state && n >= 1 && n <= kMaxPeers && rank >= 0 && rank < n,
                 "vgpt_peer_barrier: bad arguments (n=%d rank=%d, at most %d peers)", n, rank, kMaxPeers);
  PeerPtrs f;
  for (int i = 0; i < kMaxPeers; ++i) f.p[i] = i < n ? flag_ptrs[i] : nullptr;
  for (int i = 0; i < n; ++i) VGPT_CHECK_ARG(f.p[i], "vgpt_peer_barrier: null flag pointer for rank %d", i);
  peer_barrier_kernel<<<1, 32, 0, s>>>(f, n, rank, state, 20000000000ll /* ~10 s */);
  VGPT_CHECK_LAUNCH();
  return 0;
}

}  // namespace vgpt
