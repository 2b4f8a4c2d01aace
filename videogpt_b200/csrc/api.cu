// extern "C" entry points of libvgpt_b200.so (declared in include/vgpt_b200.h).
#include "../../include/vgpt_b200.h"
#include "vgpt_internal.h"

#include <cstdlib>

static_assert(VGPT_PAGE_TOKENS == 128, "attention.cu assumes 128-token pages");
#define S(stream) static_cast<cudaStream_t>(stream)

extern "C" {

int vgpt_abi_version(void) { return VGPT_ABI_VERSION; }
const char* vgpt_last_error(void) { return vgpt::last_error(); }

int vgpt_gemm_bf16(const void* A, const void* W, void* C, const void* R, int M, int N, int K, int lda,
                   int ldc, int epilogue, int block_n, int tail_mode, void* stream) {
  return vgpt::gemm_bf16(A, W, C, R, M, N, K, lda, ldc, epilogue, block_n, tail_mode, S(stream));
}
int vgpt_pack_gate_up(const void* w, void* packed, int I, int K, void* stream) {
  return vgpt::pack_gate_up(w, packed, I, K, S(stream));
}
int vgpt_rmsnorm(const void* x, const void* weight, void* y, int rows, int hidden, float eps, void* stream) {
  return vgpt::rmsnorm(x, weight, y, rows, hidden, eps, S(stream));
}
int vgpt_rope_table(const float* inv_freq, void* table, int max_pos, int head_dim, void* stream) {
  return vgpt::rope_table(inv_freq, table, max_pos, head_dim, S(stream));
}
int vgpt_rope_kv_append(void* qkv, const int32_t* row_pos, const int32_t* row_slot, const void* table,
                        void* k_pool, void* v_pool, int rows, int H, int D, void* stream) {
  void* kp[1] = {k_pool};
  void* vp[1] = {v_pool};
  return vgpt::rope_kv_append(qkv, row_pos, row_slot, table, kp, vp, 1, rows, H, D, VGPT_PAGE_TOKENS,
                              S(stream));
}
int vgpt_rope_kv_append_peers(void* qkv, const int32_t* row_pos, const int32_t* row_slot, const void* table,
                              void* const* k_pools, void* const* v_pools, int n_pools, int rows, int H,
                              int D, void* stream) {
  return vgpt::rope_kv_append(qkv, row_pos, row_slot, table, k_pools, v_pools, n_pools, rows, H, D,
                              VGPT_PAGE_TOKENS, S(stream));
}
int vgpt_final_layer_rows(const void* hidden, int rows, int hidden_size, const void* norm_weight, float rms_eps,
                          const int32_t* row_kind, const int32_t* row_a, const int32_t* row_b, const void* mod,
                          const void* w, const void* bias, void* const* preds, int n_preds, int channels, int lat_h,
                          int lat_w, void* stream) {
  return vgpt::final_layer_rows(hidden, rows, hidden_size, norm_weight, rms_eps, row_kind, row_a, row_b, mod, w, bias,
                                preds, n_preds, channels, lat_h, lat_w, S(stream));
}
int vgpt_peer_alloc(void** out, uint64_t bytes) { return vgpt::peer_alloc(out, bytes); }
int vgpt_peer_free(void* p) { return vgpt::peer_free(p); }
int vgpt_peer_export(void* p, void* handle64) { return vgpt::peer_export(p, handle64); }
int vgpt_peer_import(const void* handle64, void** out) { return vgpt::peer_import(handle64, out); }
int vgpt_peer_close(void* p) { return vgpt::peer_close(p); }
int vgpt_peer_barrier(void* const* flag_ptrs, int n_ranks, int rank, uint32_t* state, void* stream) {
  return vgpt::peer_barrier(flag_ptrs, n_ranks, rank, state, S(stream));
}
int vgpt_attn_clip_causal(const void* q, int q_ld, int q_rows, void* out, int out_ld, const void* k_pool,
                          const void* v_pool, int total_pages, const int32_t* page_table, int max_pages,
                          const VgptAttnSeq* seqs, int num_seqs, int max_q_rows, const int32_t* q_code,
                          const int32_t* k_code, const int32_t* k_tile_minmax, int max_k_tiles, int H,
                          int D, float scale, void* stream) {
  return vgpt::attn_clip_causal_pair(q, q_ld, q_rows, out, out_ld, k_pool, v_pool, total_pages, page_table,
                                     max_pages, seqs, num_seqs, max_q_rows, q_code, k_code, k_tile_minmax,
                                     max_k_tiles, H, D, scale, S(stream));
}
int vgpt_attn_clip_causal_mma_sync(const void* q, int q_ld, int q_rows, void* out, int out_ld,
                                   const void* k_pool, const void* v_pool, int total_pages,
                                   const int32_t* page_table, int max_pages, const VgptAttnSeq* seqs,
                                   int num_seqs, int max_q_rows, const int32_t* q_code, const int32_t* k_code,
                                   const int32_t* k_tile_minmax, int max_k_tiles, int H, int D, float scale,
                                   void* stream) {
  (void)q_rows; (void)total_pages;
  return vgpt::attn_clip_causal(q, q_ld, out, out_ld, k_pool, v_pool, page_table, max_pages, seqs,
                                num_seqs, max_q_rows, q_code, k_code, k_tile_minmax, max_k_tiles, H, D,
                                scale, S(stream));
}
int vgpt_embed_assemble(void* hidden, int rows, int hidden_size, const int32_t* row_kind,
                        const int32_t* row_a, const int32_t* row_b, const void* embed_tokens,
                        const void* time_tokens, const void* z, const void* ctx, int channels, int lat_h,
                        int lat_w, const void* w_noisy, const void* b_noisy, const void* w_ctx,
                        const void* b_ctx, const void* pos_rows, void* stream) {
  return vgpt::embed_assemble(hidden, rows, hidden_size, row_kind, row_a, row_b, embed_tokens,
                              time_tokens, z, ctx, channels, lat_h, lat_w, w_noisy, b_noisy, w_ctx,
                              b_ctx, pos_rows, S(stream));
}
int vgpt_timestep_sinusoid(const float* t, const float* freqs, void* out, int n, int dim, void* stream) {
  return vgpt::timestep_sinusoid(t, freqs, out, n, dim, S(stream));
}
int vgpt_linear_small(const void* in, const void* W, const void* bias, void* out, int n, int N, int K,
                      int pre_silu, int post_silu, void* stream) {
  return vgpt::linear_small(in, W, bias, out, n, N, K, pre_silu, post_silu, S(stream));
}
int vgpt_final_layer(const void* hidden, int hidden_size, const void* norm_weight, float rms_eps,
                     const int32_t* lat_row0, const void* mod, const void* w, const void* bias, void* pred,
                     int n_lat, int channels, int lat_h, int lat_w, void* z_euler, void* vel_out,
                     const float* scalars_dev, int use_cfg, int x1_mode, void* stream) {
  return vgpt::final_layer(hidden, hidden_size, norm_weight, rms_eps, lat_row0, mod, w, bias, pred, n_lat, channels,
                           lat_h, lat_w, z_euler, vel_out, scalars_dev, use_cfg, x1_mode, S(stream));
}
int vgpt_cfg_euler(void* z, const void* pred, void* vel_out, int half_numel, int use_cfg, int x1_mode,
                   float one_minus_sigma, float dsigma, float guidance, const float* scalars_dev,
                   void* stream) {
  return vgpt::cfg_euler(z, pred, vel_out, half_numel, use_cfg, x1_mode, one_minus_sigma, dsigma,
                         guidance, scalars_dev, S(stream));
}
int vgpt_cfg_combine(void* pred, int half_numel, float guidance, void* stream) {
  return vgpt::cfg_combine(pred, half_numel, guidance, S(stream));
}
int vgpt_mask_from_codes(const int32_t* q_code, const int32_t* k_code, void* out, int Lq, int Lk,
                         void* stream) {
  return vgpt::mask_from_codes(q_code, k_code, out, Lq, Lk, S(stream));
}
int vgpt_debug_attn_trace(void* out, int max_events, int* n_events, void* stream) {
  return vgpt::attn_trace_read(out, max_events, n_events, S(stream));
}

}  // extern "C"
