/* vgpt_b200 -- C ABI of the B200-native next-clip denoising path of Video-GPT.
 *
 * The reference (zhuangshaobin/Video-GPT) is pure Python and has no FFI layer; the interfaces it
 * swaps implementations through are the Python seams S1-S3 of SURVEY.md section 8(b).  Each entry
 * point below is what the Python shim for one of those seams binds (videogpt_b200/_lib.py via
 * ctypes; the stub is shown in INTEGRATION.md) and names the reference code it replaces.
 *
 * Conventions: plain pointers and sizes, no torch types.  All data pointers are DEVICE pointers
 * (bf16 = uint16 storage unless stated) in the current CUDA context; `stream` is a cudaStream_t
 * passed as void*.  Functions only enqueue work: they never allocate or free caller memory and
 * never synchronise the stream, so they are CUDA-graph capturable and re-entrant per
 * (device, stream).  Return value: 0 = ok, negative = argument error, positive = cudaError_t;
 * vgpt_last_error() returns a thread-local message for the last non-zero return.
 */
#ifndef VGPT_B200_H_
#define VGPT_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VGPT_ABI_VERSION 1
#define VGPT_PAGE_TOKENS 128 /* tokens per KV-cache page */

enum { VGPT_EPI_STORE = 0, VGPT_EPI_RESIDUAL = 1, VGPT_EPI_SWIGLU = 2 };
enum { VGPT_ROW_TOKEN = 0, VGPT_ROW_TIME = 1, VGPT_ROW_NOISY_PATCH = 2, VGPT_ROW_CONTEXT_PATCH = 3 };

/* One sequence of an attention call (device-resident array). */
typedef struct VgptAttnSeq {
  int32_t q_row0;   /* first row of this sequence's queries in q / out / q_code            */
  int32_t n_q;      /* number of query rows                                              */
  int32_t kv_len;   /* number of valid keys: logical cache positions [0, kv_len)         */
  int32_t reserved;
} VgptAttnSeq;

int vgpt_abi_version(void);
const char* vgpt_last_error(void);

/* C[M,N] = A[M,K] x W[N,K]^T, bf16 in / fp32 accumulate / bf16 out, tcgen05 + TMEM + TMA.
 * Replaces qkv_proj / o_proj (LVM/transform/sdpa_transform.py:39, 89) and gate_up_proj /
 * down_proj of transformers' Phi3MLP.  epilogue: VGPT_EPI_STORE; VGPT_EPI_RESIDUAL
 * (C = bf16(A W^T) + R, R may alias C: the in-place residual stream of Phi3DecoderLayer);
 * VGPT_EPI_SWIGLU (W packed by vgpt_pack_gate_up, C[M,N/2] = up * silu(gate)).
 * K % 64 == 0, N % 64 == 0.  CTA pairs (tcgen05 cta_group::2) compute 256 x block_n tiles, block_n 128 / 192 / 256,
 * 0 = tuned default.  tail_mode: when M = 256 q + t with q >= 1 and 0 < t <= 32, the t tail rows can be computed
 * inside the k-loop of the last full tile row (operands swapped, from the W k-blocks already in shared memory)
 * instead of by a (q+1)-th, almost empty row of tiles: 3 = do that, 1 = plain tiles only, -1 = tuned default.
 * A row gets the same bits either way. */
int vgpt_gemm_bf16(const void* A, const void* W, void* C, const void* R, int M, int N, int K, int lda,
                   int ldc, int epilogue, int block_n, int tail_mode, void* stream);

/* gate_up_proj.weight [2I,K] ([gate | up] rows, Phi3MLP chunk(2)) -> block-interleaved rows. */
int vgpt_pack_gate_up(const void* w, void* packed, int I, int K, void* stream);

/* Phi3RMSNorm (transformers 4.47.1): y = w * bf16(x * rsqrt(mean(x^2) + eps)). */
int vgpt_rmsnorm(const void* x, const void* weight, void* y, int rows, int hidden, float eps, void* stream);

/* cos/sin table [max_pos][head_dim] bf16 = cos[0:D/2] | sin[0:D/2] of pos * inv_freq (fp32)
 * (Phi3RotaryEmbedding.forward, called at sdpa_transform.py:52). */
int vgpt_rope_table(const float* inv_freq, void* table, int max_pos, int head_dim, void* stream);

/* apply_rotary_pos_emb on q (in place in qkv[rows, 3*H*D]) and k, and the KV-cache update
 * (sdpa_transform.py:52-57): post-RoPE k and v of each row go to slot row_slot[row]
 * (= pool_page * VGPT_PAGE_TOKENS + offset; < 0: not cached) of pools [page][H][128][D]. */
int vgpt_rope_kv_append(void* qkv, const int32_t* row_pos, const int32_t* row_slot, const void* table,
                        void* k_pool, void* v_pool, int rows, int H, int D, void* stream);

/* Sequence-parallel form of vgpt_rope_kv_append: the post-RoPE k / v rows are stored into n_pools
 * pools (this rank's and every peer's, mapped with vgpt_peer_import) -- the K/V all-gather of the
 * sequence-parallel path fused into the producing kernel as NVLink stores.  Replaces the four
 * Ulysses all-to-alls per layer of LVM/transform/sdpa_transform.py:126-156.  n_pools <= 8. */
int vgpt_rope_kv_append_peers(void* qkv, const int32_t* row_pos, const int32_t* row_slot, const void* table,
                              void* const* k_pools, void* const* v_pools, int n_pools, int rows, int H,
                              int D, void* stream);

/* Clip-block-causal flash attention over the paged KV pools (tcgen05 + TMEM + TMA); replaces
 * F.scaled_dot_product_attention with the dense additive mask (sdpa_transform.py:152-166,
 * OmniGen/transformer.py:128-145).  allowed(q,k) <=> q_code[q] >= k_code[k]; k_code is
 * [num_seqs][max_pages*128], k_tile_minmax [num_seqs][max_k_tiles][2] holds (min,max) of k_code
 * per 64-key tile, page_table [num_seqs][max_pages] indexes pools of total_pages pages.  q/out
 * rows are [.., H*D] with leading dimensions q_ld / out_ld (q may point into the fused qkv
 * buffer of q_rows rows). */
int vgpt_attn_clip_causal(const void* q, int q_ld, int q_rows, void* out, int out_ld, const void* k_pool,
                          const void* v_pool, int total_pages, const int32_t* page_table, int max_pages,
                          const VgptAttnSeq* seqs, int num_seqs, int max_q_rows, const int32_t* q_code,
                          const int32_t* k_code, const int32_t* k_tile_minmax, int max_k_tiles, int H,
                          int D, float scale, void* stream);

/* Same contract on the legacy tensor path (mma.sync, cp.async): the round-1 kernel, kept only as
 * an independent cross-check of the tcgen05 kernel in the GPU tests. */
int vgpt_attn_clip_causal_mma_sync(const void* q, int q_ld, int q_rows, void* out, int out_ld,
                                   const void* k_pool, const void* v_pool, int total_pages,
                                   const int32_t* page_table, int max_pages, const VgptAttnSeq* seqs,
                                   int num_seqs, int max_q_rows, const int32_t* q_code, const int32_t* k_code,
                                   const int32_t* k_tile_minmax, int max_k_tiles, int H, int D, float scale,
                                   void* stream);

/* Sequence assembly (LVM/model.py:419-454): hidden[row] = embed_tokens[a] | time_tokens[a] |
 * PatchEmbedMR(z[a] or ctx[a]) patch b + pos_rows[b], per row_kind (VGPT_ROW_*).
 * z / ctx: [n, C=4, lat_h, lat_w]; conv weights [hidden, 4, 2, 2]; pos_rows [tokens, hidden]. */
int vgpt_embed_assemble(void* hidden, int rows, int hidden_size, const int32_t* row_kind,
                        const int32_t* row_a, const int32_t* row_b, const void* embed_tokens,
                        const void* time_tokens, const void* z, const void* ctx, int channels, int lat_h,
                        int lat_w, const void* w_noisy, const void* b_noisy, const void* w_ctx,
                        const void* b_ctx, const void* pos_rows, void* stream);

/* TimestepEmbedder.timestep_embedding (LVM/model.py:39-58): out[n, dim] = [cos | sin](t * freqs). */
int vgpt_timestep_sinusoid(const float* t, const float* freqs, void* out, int n, int dim, void* stream);

/* out[n,N] = post(pre(in[n,K]) W[N,K]^T + bias), n <= 16, optional SiLU before / after
 * (TimestepEmbedder.mlp and FinalLayer.adaLN_modulation, LVM/model.py:32-36, 74-77). */
int vgpt_linear_small(const void* in, const void* W, const void* bias, void* out, int n, int N, int K,
                      int pre_silu, int post_silu, void* stream);

/* [llm.norm ->] FinalLayer + unpatchify [-> scheduler update] (OmniGen/transformer.py:214, LVM/model.py:79-83,
 * 255-265, 478-486, LVM/scheduler.py:178-204): for latent j the image-token rows hidden[lat_row0[j] .. +tokens)
 * -> pred[j, C, lat_h, lat_w]; mod[j] = [shift | scale].  norm_weight != NULL: `hidden` holds the RAW residual
 * stream and the final Phi3RMSNorm (weight norm_weight, eps rms_eps) is applied first.  z_euler != NULL: the step's
 * x1 -> v / CFG / Euler update of vgpt_cfg_euler is applied to z_euler right behind the prediction (latents laid out
 * [cond | uncond], scalars_dev = device float[3] {1 - sigma, d sigma, guidance}; vel_out optional) -- one launch
 * instead of three at the end of every Euler step, same bits. */
int vgpt_final_layer(const void* hidden, int hidden_size, const void* norm_weight, float rms_eps,
                     const int32_t* lat_row0, const void* mod, const void* w, const void* bias, void* pred,
                     int n_lat, int channels, int lat_h, int lat_w, void* z_euler, void* vel_out,
                     const float* scalars_dev, int use_cfg, int x1_mode, void* stream);

/* Row-driven FinalLayer + unpatchify for row-sharded plans: local row r is an image token of
 * latent row_a[r], patch row_b[r] when row_kind[r] == VGPT_ROW_NOISY_PATCH (other rows are
 * skipped); the 16 outputs per row are stored into every destination preds[0..n_preds) (own +
 * peers), replacing the hidden-state all-gather of LVM/model.py:466-474.  norm_weight as in vgpt_final_layer. */
int vgpt_final_layer_rows(const void* hidden, int rows, int hidden_size, const void* norm_weight, float rms_eps,
                          const int32_t* row_kind, const int32_t* row_a, const int32_t* row_b, const void* mod,
                          const void* w, const void* bias, void* const* preds, int n_preds, int channels, int lat_h,
                          int lat_w, void* stream);

/* Peer memory for the sequence-parallel path (one process per GPU, NVLink/NVSwitch).  These five
 * are resource management, not data path: vgpt_peer_alloc = zero-filled cudaMalloc (exportable),
 * vgpt_peer_export writes the 64-byte CUDA IPC handle of an allocation, vgpt_peer_import maps a
 * peer's allocation into this process (peer access enabled lazily), vgpt_peer_close / _free undo
 * them.  The reference gets its process groups from torch.distributed / deepspeed
 * (LVM/acceleration/parallel_states.py:25-60). */
int vgpt_peer_alloc(void** out, uint64_t bytes);
int vgpt_peer_free(void* p);
int vgpt_peer_export(void* p, void* handle64);
int vgpt_peer_import(const void* handle64, void** out);
int vgpt_peer_close(void* p);

/* Barrier over n_ranks GPUs through peer-mapped flag words: flag_ptrs[j] -> rank j's uint32[n_ranks]
 * (flag_ptrs[rank] is local); state = local, 8-byte aligned uint32[4] {epoch, timed_out, uint64 nanoseconds
 * spent inside barrier kernels so far}.  Enqueued like any other
 * kernel (graph capturable); every rank must enqueue the same sequence of barriers.  A peer that
 * never arrives sets state[1] after ~10 s instead of hanging the GPU; later barriers then return
 * at once (fail fast; the host checks state[1]). */
int vgpt_peer_barrier(void* const* flag_ptrs, int n_ranks, int rank, uint32_t* state, void* stream);

/* x1 -> velocity, CFG and the Euler update (LVM/scheduler.py:178-204, LVM/model.py:554-562) on
 * z/pred laid out [cond latents | uncond latents] (half_numel elements each; one half if
 * !use_cfg).  scalars_dev (optional, device float[3] = {1-sigma, sigma_next-sigma, guidance})
 * overrides the by-value scalars so a captured graph can be replayed per step.
 * vel_out (optional) receives the applied velocity of the cond half. */
int vgpt_cfg_euler(void* z, const void* pred, void* vel_out, int half_numel, int use_cfg, int x1_mode,
                   float one_minus_sigma, float dsigma, float guidance, const float* scalars_dev,
                   void* stream);

/* v-mode CFG inside the model (LVM/model.py:554-562): both halves of pred <- u + g * (c - u). */
int vgpt_cfg_combine(void* pred, int half_numel, float guidance, void* stream);

/* out[q][k] = q_code[q] >= k_code[k] (uint8): the reference's dense mask
 * (LVM/processor.py:682-731) from the codes the attention kernel consumes. */
int vgpt_mask_from_codes(const int32_t* q_code, const int32_t* k_code, void* out, int Lq, int Lk,
                         void* stream);

/* Diagnostic: with VGPT_ATTN_VARIANT=8 the attention kernel's CTA (0, 0, 0) records a clock64 time stamp
 * at every hand-over between its roles (TMA producer, MMA issuer, softmax warps).  Copies the events
 * recorded so far into out[2 * max_events] uint64 (host memory; pairs of clock and
 * warp << 40 | tile << 32 | kv_tile << 8 | event), resets the recorder, synchronises the stream.
 * Not on any product path (tools/attn_trace.py). */
int vgpt_debug_attn_trace(void* out, int max_events, int* n_events, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* VGPT_B200_H_ */
