/*
 * ls_gemm.h -- C ABI of the sm_100a tensor-core GEMM used by the encoder / decoder layers.
 *
 * What it replaces in the reference (Chrixtar/latentsplat): the cuBLAS calls behind every
 * `nn.Linear` on the hot path -- `to_q/to_kv/to_out` (src/model/transformer/attention.py:45-52),
 * the token MLPs (feed_forward.py:31-37, backbone_dino.py:34-43), `backbone_projection`, `to_gaussians`,
 * the depth head (encoder_epipolar.py:70-73, 97-103; depth_predictor_monocular.py:25-28), the DINO ViT
 * linears -- forward, input-gradient and weight-gradient GEMMs.  The reference runs them as fp32 SIMT
 * GEMMs (torch default matmul precision "highest"); here they are tcgen05.mma kind::tf32 with fp32
 * operands read straight from HBM by TMA (no cast passes), fp32 accumulation in TMEM, and bias /
 * activation fused into the epilogue.
 *
 *   C[M,N] (+)= act(A[M,K] * B[N,K]^T (+ bias[N])) (+ residual[M,N])
 *
 * Either operand may be K-major (row-major [rows][K], e.g. activations X and weights W in Y = X W^T)
 * or MN-major (stored [K][rows]; e.g. W in dX = dY W, and both dY and X in dW = dY^T X), so forward,
 * dgrad and wgrad are the same kernel without any transposed copies.
 *
 * Conventions as ls_raster.h: device pointers, caller-owned buffers, work enqueued on `stream`, no sync,
 * 0 / negative return + ls_last_error().  Pointers must be 16-byte aligned and leading dimensions
 * multiples of 4 elements (TMA global-stride rule).
 */
#ifndef LS_GEMM_H
#define LS_GEMM_H

#include <stdint.h>

#include "ls_raster.h" /* LS_API, ls_last_error */

#ifdef __cplusplus
extern "C" {
#endif

enum { LS_ACT_NONE = 0, LS_ACT_RELU = 1, LS_ACT_GELU = 2 /* exact erf GELU, torch.nn.GELU() */, LS_ACT_SILU = 3,
       LS_ACT_LRELU = 4 /* LeakyReLU(0.2), discriminator_patch_gan.py:24 */ };

typedef struct LsGemmArgs {
    int32_t M, N, K;
    int32_t a_mn_major;  /* 0: A[m*lda + k] (K-major)   1: A[k*lda + m] (MN-major) */
    int32_t b_mn_major;  /* 0: B[n*ldb + k] (K-major)   1: B[k*ldb + n] (MN-major) */
    int32_t act;         /* LS_ACT_*; must be NONE when split_k > 1 or accumulate != 0 */
    int32_t split_k;     /* >= 1: slices of the K loop, partial tiles combined with red.global.add;
                            the caller zero-fills C first (or passes accumulate=1 semantics)      */
    int32_t accumulate;  /* 1: C += result (atomic adds), 0: C = result                           */
    int64_t lda, ldb, ldc;
    const float* A;
    const float* B;
    float* C;            /* [M][ldc] row-major fp32                                               */
    const float* bias;   /* (N) or NULL                                                           */
    const float* residual; /* [M][ldr] or NULL: C = act(A B^T + bias) + residual -- the skip connection of a
                              transformer block (x + proj(...), x + fc2(...)) in the epilogue; not with split-K  */
    int64_t ldr;
    float* pre_out;      /* [M][ldc] or NULL: the pre-activation A B^T + bias, kept for the backward of GELU     */
} LsGemmArgs;

LS_API int ls_gemm_tf32(const LsGemmArgs* args, void* stream /* cudaStream_t */);

/* Fused single-query multi-head attention: replaces the chunk / rearrange / bmm / softmax / bmm sequence of
 * src/model/transformer/attention.py:54-70 for the epipolar cross-attention (one query per ray, S <= 32 sampled
 * key/value tokens, epipolar_transformer.py:127-135).  q (R, H*D), kv (R, S, 2*H*D) = [K | V] as produced by to_kv,
 * out (R, H*D), p (R, H, S) softmax probabilities kept for backward.  D must be 128. */
LS_API int ls_sq_attention_forward(const float* q, const float* kv, float* out, float* p, int32_t R, int32_t H, int32_t S,
                                   int32_t D, float scale, void* stream);
LS_API int ls_sq_attention_backward(const float* q, const float* kv, const float* p, const float* dout, float* dq,
                                    float* dkv, int32_t R, int32_t H, int32_t S, int32_t D, float scale, void* stream);

/* Weight-absorbed form of the same attention: with one query per ray and bias-free to_q / to_kv,
 *   score = (W_k,h^T q_h) . z_j   and   out_h = W_v,h (sum_j p_{h,j} z_j),
 * so the kv tensor (S * 2*H*D floats per ray) is never materialised.  qt (R, H, 128) absorbed queries, z (R, S, 128) raw
 * samples, zbar (R, H, 128) probability-weighted sample means, p (R, H, S).  The per-head 128x128 GEMMs around these
 * kernels (q -> qt, zbar -> out, and their gradients) are ls_gemm_tf32 calls made by the caller. */
LS_API int ls_absorbed_attention_forward(const float* qt, const float* z, float* zbar, float* p, int32_t R, int32_t H,
                                         int32_t S, int32_t Dz, float scale, void* stream);
LS_API int ls_absorbed_attention_backward(const float* qt, const float* z, const float* p, const float* dzbar, float* dqt,
                                          float* dz, int32_t R, int32_t H, int32_t S, int32_t Dz, float scale, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* LS_GEMM_H */
