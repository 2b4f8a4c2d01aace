// Internal (C++) declarations shared by the .cu translation units; the public C ABI is
// include/vgpt_b200.h.
#pragma once

#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace vgpt {

void set_last_error(const char* fmt, ...);
int device_sm_count();

// Peer memory (sequence-parallel path): at most kMaxPeers ranks share one video; a kernel that
// produces data every rank needs takes the destination of each rank by value.
constexpr int kMaxPeers = 8;
struct PeerPtrs { void* p[kMaxPeers]; };

// cuTensorMapEncodeTiled through the runtime's driver entry point (no link-time libcuda).
int encode_tensor_map(CUtensorMap* map, CUtensorMapDataType dtype, uint32_t rank, void* base,
                      const cuuint64_t* dims, const cuuint64_t* strides_bytes, const cuuint32_t* box,
                      const cuuint32_t* elem_strides, CUtensorMapSwizzle swizzle);

const char* last_error();

int gemm_bf16(const void* A, const void* W, void* C, const void* R, int M, int N, int K, int lda,
              int ldc, int epilogue, int block_n, int tail_mode, cudaStream_t stream);
int debug_gemm_flags();
int pick_pair_block_n(int M, int N, int num_sms);
int pack_gate_up(const void* w, void* packed, int I, int K, cudaStream_t s);
int rmsnorm(const void* x, const void* w, void* y, int rows, int hidden, float eps, cudaStream_t s);
int rope_table(const float* inv_freq, void* tab, int max_pos, int head_dim, cudaStream_t s);
int rope_kv_append(void* qkv, const int32_t* row_pos, const int32_t* row_slot, const void* tab,
                   void* const* k_pools, void* const* v_pools, int n_pools, int rows, int H, int D,
                   int page_tokens, cudaStream_t s);
int attn_clip_causal(const void* q, int q_ld, void* out, int out_ld, const void* k_pool,
                     const void* v_pool, const int32_t* page_table, int max_pages, const void* seqs,
                     int num_seqs, int max_q_rows, const int32_t* q_code, const int32_t* k_code,
                     const int32_t* k_tile_minmax, int max_k_tiles, int H, int D, float scale,
                     cudaStream_t s);
int attn_clip_causal_pair(const void* q, int q_ld, int q_rows, void* out, int out_ld, const void* k_pool,
                          const void* v_pool, int total_pages, const int32_t* page_table, int max_pages,
                          const void* seqs, int num_seqs, int max_q_rows, const int32_t* q_code,
                          const int32_t* k_code, const int32_t* k_tile_minmax, int max_k_tiles, int H, int D,
                          float scale, cudaStream_t s);
int embed_assemble(void* hidden, int rows, int hs, const int32_t* kind, const int32_t* a,
                   const int32_t* b, const void* embed_tokens, const void* time_tokens, const void* z,
                   const void* ctx, int C, int lat_h, int lat_w, const void* wx, const void* bx,
                   const void* wc, const void* bc, const void* pos, cudaStream_t s);
int timestep_sinusoid(const float* t, const float* freqs, void* out, int n, int dim, cudaStream_t s);
int linear_small(const void* in, const void* W, const void* bias, void* out, int n, int N, int K,
                 int pre_silu, int post_silu, cudaStream_t s);
int final_layer(const void* hidden, int hs, const void* norm_w, float rms_eps, const int32_t* lat_row0, const void* mod,
                const void* w, const void* bias, void* pred, int n_lat, int C, int lat_h, int lat_w, void* z_euler,
                void* vel_out, const float* scalars_dev, int use_cfg, int x1_mode, cudaStream_t s);
int final_layer_rows(const void* hidden, int rows, int hs, const void* norm_w, float rms_eps, const int32_t* kind,
                     const int32_t* a, const int32_t* b, const void* mod, const void* w, const void* bias,
                     void* const* preds, int n_preds, int C, int lat_h, int lat_w, cudaStream_t s);
int peer_alloc(void** out, uint64_t bytes);
int peer_free(void* p);
int peer_export(void* p, void* handle64);
int peer_import(const void* handle64, void** out);
int peer_close(void* p);
int peer_barrier(void* const* flag_ptrs, int n, int rank, uint32_t* state, cudaStream_t s);
int cfg_euler(void* z, const void* pred, void* vel_out, int half_numel, int use_cfg, int x1_mode,
              float one_minus_sigma, float dsigma, float guidance, const float* scalars_dev,
              cudaStream_t s);
int cfg_combine(void* pred, int half_numel, float guidance, cudaStream_t s);
int mask_from_codes(const int32_t* qc, const int32_t* kc, void* out, int Lq, int Lk, cudaStream_t s);
int attn_trace_read(void* out, int max_events, int* n_events, cudaStream_t s);

}  // namespace vgpt
