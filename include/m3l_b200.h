/* m3l_b200 — C-ABI of the B200-native VTMAE hot path.
 *
 * Plain C: pointers, sizes and a CUDA stream handle (void* == cudaStream_t); no torch types.
 * All device pointers must be valid on the current CUDA device.  Every entry point returns 0 on
 * success or a non-zero status (1 invalid argument, 2 CUDA failure, 3 unsupported shape);
 * m3l_last_error() then holds a message for the calling thread.  Nothing here synchronises the
 * stream or allocates device memory, so every call can be captured in a CUDA graph.
 *
 * Each entry point names the reference interface it stands in for.  The reference is pure
 * PyTorch (no FFI of its own: SURVEY.md §2.1), so "replaces" means the eager ops that
 * /root/reference/models/pretrain_models.py dispatches at the cited lines; INTEGRATION.md shows
 * the ctypes binding a maintainer of the reference would add.
 */
#ifndef M3L_B200_H_
#define M3L_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Message of the last failing call on this thread ("" if none). */
const char* m3l_last_error(void);

/* ------------------------------------------------------------------------------------------
 * Dense contraction with fused epilogue (tcgen05 / TMEM / TMA):
 *     out[m, n] = epilogue(alpha * sum_k A[m, k] * B[n, k])
 * Replaces nn.Linear forward / dgrad / wgrad as dispatched by vit_pytorch's Attention and
 * FeedForward (pretrain_models.py:113,784), the patch-embedding Linear (:769-771,776-778) and
 * the to_pixels / to_tactiles heads (:115-116,328,333).
 * ---------------------------------------------------------------------------------------- */
typedef struct m3l_gemm_args {
  const void* a;      /* bf16. K-major: [m, k] row-major (lda). MN-major: [k, m] row-major (lda) */
  const void* b;      /* bf16. K-major: [n, k] row-major (ldb). MN-major: [k, n] row-major (ldb) */
  int32_t lda, ldb;   /* leading dimensions in elements, multiples of 8 */
  int32_t a_mn_major; /* 0: K-major, 1: MN-major (must equal b_mn_major) */
  int32_t b_mn_major;
  int32_t m, n, k;    /* n a multiple of 8 */
  int32_t splits;     /* split-K factor; > 1 requires out_mode 2 */
  int32_t bn;         /* N tile: 0 auto, 64, 128 or 256 */
  void* out;          /* [m, ldo] */
  int32_t ldo;
  int32_t out_mode;   /* 0: bf16 store, 1: fp32 store, 2: fp32 atomic accumulate (red.add) */
  const float* bias;  /* [n] fp32 or NULL */
  const void* residual; /* bf16 [m, ldr] or NULL; added after the activation (may alias out) */
  int32_t ldr;
  int32_t act;        /* 0 none; 1 exact-erf GELU (pre-activation stored to aux_out if non-NULL);
                         2 multiply by GELU'(aux_in[m, n]) (backward of 1) */
  void* aux_out;      /* bf16 [m, ld_aux] or NULL */
  const void* aux_in; /* bf16 [m, ld_aux] or NULL */
  int32_t ld_aux;
  float alpha;
} m3l_gemm_args;

int m3l_gemm_bf16(const m3l_gemm_args* args, void* stream);

#ifdef __cplusplus
}
#endif

#endif /* M3L_B200_H_ */
