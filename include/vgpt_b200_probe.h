/* libvgpt_b200_probe.so -- tcgen05 descriptor / issue-rate probes (videogpt_b200/csrc/probe/).
 *
 * Test and tuning hooks only: tests/test_umma_layouts.py pins the shared-memory layouts and descriptor encodings the
 * production kernels rely on, tools/umma_rate.py measures cycles per tcgen05.mma.  Deliberately NOT part of the
 * product library libvgpt_b200.so (include/vgpt_b200.h).  Same conventions: int return (0 ok, < 0 argument error,
 * > 0 cudaError_t), work enqueued on `stream`, vgpt_probe_last_error() for the message. */
#ifndef VGPT_B200_PROBE_H
#define VGPT_B200_PROBE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

const char* vgpt_probe_last_error(void);

/* Test hook: run k_steps tcgen05.mma on raw shared-memory images with caller-built descriptors. */
int vgpt_debug_umma_probe(const void* a_img, int a_bytes, const void* b_img, int b_bytes,
                          uint64_t a_desc_base, uint64_t b_desc_base, uint32_t idesc, int k_steps,
                          uint32_t a_step_bytes, uint32_t b_step_bytes, float* d_out, int n_cols,
                          void* stream);

/* Test hook: same with the A operand in tensor memory (a_words[128][a_cols] packed bf16x2). */
int vgpt_debug_umma_probe_ts(const void* a_words, int a_cols, const void* b_img, int b_bytes,
                             uint64_t b_desc_base, uint32_t idesc, int k_steps, uint32_t b_step_bytes,
                             float* d_out, int n_cols, void* stream);

/* Test / tuning hook: cycles per back-to-back tcgen05.mma (M = 128, cta_group::1, bf16) of width N.
 * mode 0 = SS K-major SW128, 1 = SS K-major SW64, 2 = TS + MN-major SW128 B, 3 = TS + MN-major SW64 B,
 * 4 = CTA pairs (cta_group::2, M = 256, SS K-major SW128; `ctas` = clusters; not yet run on hardware);
 * n_acc = 1 dependent chain, 2 alternating accumulators; commit_every = 0 / 1 / 2 / 4 / 8: a tcgen05.commit
 * after every that many MMAs; out[ctas] = cycles per MMA per CTA. */
int vgpt_debug_umma_rate(int mode, int N, int iters, int n_acc, int commit_every, int ctas, float* out,
                         void* stream);

#ifdef __cplusplus
}
#endif
#endif /* VGPT_B200_PROBE_H */
