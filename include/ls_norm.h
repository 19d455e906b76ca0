/*
 * ls_norm.h -- C ABI of the fused GroupNorm (+ SiLU) of the VAE decoder, NCHW fp32.
 *
 * What it replaces in the reference (Chrixtar/latentsplat): the `GroupNorm(32, C, eps=1e-6) -> SiLU` pairs of the
 * diffusers VAE decoder the reference wraps (src/model/autoencoder/autoencoder_kl.py:93-124: ResnetBlock2D.norm1/norm2 +
 * nonlinearity, Decoder.conv_norm_out + conv_act) and the activation-free GroupNorm of the mid-block attention.
 *
 *   y = act( (x - mean_g) * rstd_g * gamma_c + beta_c ),   statistics over (C/G channels x H x W) per image
 *
 * Conventions as ls_raster.h: device pointers, caller-owned buffers, work enqueued on `stream`, no sync, 0 / negative
 * return + ls_last_error().  x / y / dy / dx 16-byte aligned, H*W a multiple of 4.
 */
#ifndef LS_NORM_H
#define LS_NORM_H

#include <stdint.h>

#include "ls_raster.h" /* LS_API, ls_last_error */

#ifdef __cplusplus
extern "C" {
#endif

typedef struct LsGroupNorm {
    int32_t N, C, G;      /* images, channels, groups (C % G == 0)                                             */
    int32_t act;          /* 0: none, 1: SiLU                                                                  */
    int64_t HW;           /* H * W                                                                             */
    float eps;
    const float* x;       /* (N, C, H, W)                                                                      */
    const float* gamma;   /* (C)                                                                               */
    const float* beta;    /* (C)                                                                               */
    double* stats;        /* (N, G, 2): sum, sum of squares -- written by forward, read by backward            */
} LsGroupNorm;

LS_API int ls_groupnorm_forward(const LsGroupNorm* args, float* y, void* stream /* cudaStream_t */);
/* dy (N,C,H,W) -> dx (N,C,H,W); sums (N, C, 2) receives per-(image, channel) { sum ds, sum ds*xhat } with
 * ds = dy * act'(u): d beta_c = sum_n sums[n,c,0], d gamma_c = sum_n sums[n,c,1] (left to the caller: N*C values). */
LS_API int ls_groupnorm_backward(const LsGroupNorm* args, const float* dy, float* dx, double* sums, void* stream /* cudaStream_t */);

/* The same GroupNorm (+ SiLU) for NHWC activations (N, H*W, C) -- the layout the implicit-GEMM convolutions (ls_conv.h) keep
 * the VAE decoder in.  Same argument struct, except that `stats` is (N, C, 2): per-(image, channel) {sum x, sum x^2} (group
 * statistics are assembled from them), and `sums` of the backward is (N, C, 2) as above.  C % 4 == 0, C <= 1024.          */
LS_API int ls_groupnorm_nhwc_forward(const LsGroupNorm* args, float* y, void* stream /* cudaStream_t */);
LS_API int ls_groupnorm_nhwc_backward(const LsGroupNorm* args, const float* dy, float* dx, double* sums, void* stream /* cudaStream_t */);

/* LayerNorm over the last dimension (C a multiple of 128, <= 1024), rows x C row-major fp32: nn.LayerNorm of the DINO
 * ViT blocks (the backbone behind src/model/encoder/backbone/backbone_dino.py:33) and of the epipolar transformer's
 * PreNorm (src/model/transformer/pre_norm.py:28-35).  mean_rstd (rows, 2) is written by forward and read by backward;
 * backward ACCUMULATES into dgamma / dbeta (C each; the caller zero-fills them). */
LS_API int ls_layernorm_forward(const float* x, const float* gamma, const float* beta, float* y, float* mean_rstd,
                                int64_t rows, int32_t C, float eps, void* stream /* cudaStream_t */);
LS_API int ls_layernorm_backward(const float* x, const float* dy, const float* gamma, const float* mean_rstd, float* dx,
                                 float* dgamma, float* dbeta, int64_t rows, int32_t C, void* stream /* cudaStream_t */);

/* Per-channel bias of a convolution output, NCHW fp32: `Conv2d(..., bias=True)` everywhere on the path (VAE decoder,
 * encoder refinement / feed-forward / skip convolutions, PatchGAN; e.g. src/model/encoder/epipolar/epipolar_transformer.py:
 * 66-73, 155-170).  The convolution itself (cuDNN) runs bias-free; ls_conv_bias_add adds bias[c] to y IN PLACE,
 * ls_conv_bias_grad ACCUMULATES sum_{n,h,w} dy into dbias[c] (the caller zero-fills dbias). */
LS_API int ls_conv_bias_add(float* y, const float* bias, int64_t N, int32_t C, int64_t HW, void* stream /* cudaStream_t */);
LS_API int ls_conv_bias_grad(const float* dy, float* dbias, int64_t N, int32_t C, int64_t HW, void* stream /* cudaStream_t */);

/* out[c] += sum_r x[r*ld + c] for a row-major (rows, cols) fp32 matrix: the bias gradient `dY.sum(0)` of every nn.Linear
 * on the path (see ls_gemm.h for the list).  ACCUMULATES (the caller zero-fills out); x 16-byte aligned, ld % 4 == 0. */
LS_API int ls_col_sum(const float* x, float* out, int64_t rows, int32_t cols, int64_t ld, void* stream /* cudaStream_t */);

#ifdef __cplusplus
}
#endif
#endif /* LS_NORM_H */
