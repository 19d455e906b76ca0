/*
 * ls_fmha.h -- C ABI of the sm_100a multi-head self-attention core (flash attention on tcgen05, TF32 operands, fp32
 * accumulation in TMEM): softmax(Q K^T * scale) V without materialising the score matrix.
 *
 * What it replaces in the reference (Chrixtar/latentsplat):
 *   the attention of the DINO ViT-B/8 backbone (src/model/encoder/backbone/backbone_dino.py:33, :52 ->
 *     facebookresearch/dino vision_transformer.Attention: 12 heads x 64, 1025 tokens, 12 layers),
 *   the explicit `softmax(q k^T * scale) v` of src/model/transformer/attention.py:64-68 as used by ImageSelfAttention
 *     (src/model/encoder/epipolar/image_self_attention.py:57-79: 4 heads x 128, 256 tokens),
 * which torch runs as cuBLAS batched GEMMs + a softmax kernel (reference) or a library flash kernel (our round 1, bf16).
 * Operands stay fp32 in HBM (the tensor core truncates to TF32), so nothing is narrower than the reference's own math.
 *
 * Layout: q, k, v are (B, L, ...) token-major matrices with an arbitrary row stride, heads side by side in a row -- a packed
 * qkv projection (B, L, 3, H, D) is passed as three pointers into the same buffer with ld = 3*H*D.  o is (B, L, H*D).
 * lse (B, H, L) holds the row-wise log2-sum-exp of the scaled scores (written by forward, read by backward).
 * Conventions as ls_raster.h: device pointers (16-byte aligned), caller-owned, enqueued on `stream`, 0 / negative return.
 */
#ifndef LS_FMHA_H
#define LS_FMHA_H

#include <stdint.h>

#include "ls_raster.h" /* LS_API, ls_last_error */

#ifdef __cplusplus
extern "C" {
#endif

typedef struct LsFmha {
    int32_t B, H, L, D;       /* batch, heads, tokens, head dim (64 or 128)                              */
    float scale;              /* softmax scale (D^-0.5)                                                  */
    int32_t reserved0;
    int64_t ld_q, ld_k, ld_v; /* row strides (floats) of the token-major q / k / v matrices, % 4 == 0   */
    int64_t ld_o;             /* row stride of o (and do), >= H*D                                        */
    const float* q;           /* element (b, l, h, d) at q[(b*L + l)*ld_q + h*D + d]                     */
    const float* k;
    const float* v;
    float* o;                 /* (B*L, ld_o)                                                             */
    float* lse;               /* (B, H, L) log2-domain log-sum-exp                                       */
} LsFmha;

LS_API int ls_fmha_forward(const LsFmha* a, void* stream /* cudaStream_t */);

/* dq, dk, dv: same addressing as q, k, v (ld_q, ld_k, ld_v), every element written.  delta (B, H, L) scratch. */
LS_API int ls_fmha_backward(const LsFmha* a, const float* d_o, float* dq, float* dk, float* dv, float* delta, void* stream);

/* Wide single-head attention (the VAE mid block of src/model/autoencoder/autoencoder_kl.py:99-107: 1024 tokens x 512
 * channels, one head): the score matrix is small, so it is materialised by ls_gemm_tf32 (S = Q K^T, O = P V and the four
 * gradient GEMMs, all without transposed copies) and these two kernels do the row softmax in place.
 *   forward:  x[r, :] <- softmax(scale * x[r, :])
 *   backward: dp[r, :] <- scale * p[r, :] * (dp[r, :] - sum_j p[r, j] dp[r, j])       (dS from dP)
 * rows x cols fp32, row stride ld (floats); cols <= 4096, cols % 4 == 0, ld % 4 == 0, 16-byte aligned. */
LS_API int ls_softmax_rows_forward(float* x, int64_t rows, int32_t cols, int64_t ld, float scale, void* stream);
LS_API int ls_softmax_rows_backward(const float* p, float* dp, int64_t rows, int32_t cols, int64_t ld, float scale, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* LS_FMHA_H */
