/*
 * ls_conv.h -- C ABI of the sm_100a implicit-GEMM convolutions (TMA tile loads of NHWC activations -> shared
 * memory -> tcgen05.mma kind::tf32 -> TMEM -> fused bias / activation epilogue).
 *
 * What it replaces in the reference (Chrixtar/latentsplat): every cuDNN convolution on the render hot path --
 *   7x7 convolutions of the epipolar transformer   src/model/encoder/epipolar/epipolar_transformer.py:68-74, 164-170
 *   4x4 stride-4 down / transposed up convolutions same file :68-69, image_self_attention.py:41-52
 *   7x7 high-resolution skip                       src/model/encoder/encoder_epipolar.py:106-109
 *   3x3 / 1x1 convolutions of the VAE decoder      src/model/autoencoder/autoencoder_kl.py:61-74, 93-124 (diffusers 0.25.1)
 *   4x4 stride-2 / stride-1 PatchGAN convolutions  src/model/discriminator/discriminator_patch_gan.py:28-103
 * forward, input-gradient and weight-gradient passes.  The reference runs them through torch's cuDNN binding
 * (NCHW tensors, TF32 allowed); here activations are NHWC end to end, so a convolution is a GEMM whose A tile is a
 * (channels x pixels) box of the activation tensor shifted by the filter tap -- fetched by one 4-D TMA load with
 * zero fill outside the image (= the padding) and element strides (= the convolution stride) -- and nothing is
 * ever unfolded (no im2col buffer) or transposed (no NCHW<->NHWC passes).
 *
 * Layouts (all fp32, device pointers, 16-byte aligned):
 *   activations  (N, H, W, C)  NHWC, C % 4 == 0       == a torch NCHW tensor in channels_last memory format
 *   Conv2d weight          (Cout, R, S, Cin)          == torch (Cout, Cin, R, S) in channels_last memory format
 *   ConvTranspose2d weight (Cin, R, S, Cout)          == torch (Cin, Cout, R, S) in channels_last memory format
 * Conventions as ls_raster.h: caller-owned buffers, work enqueued on `stream`, no synchronisation, 0 / negative
 * return + ls_last_error().  Nothing is allocated; weight gradients are zero-filled by the library.
 */
#ifndef LS_CONV_H
#define LS_CONV_H

#include <stdint.h>

#include "ls_gemm.h" /* LS_API, LS_ACT_* */

#ifdef __cplusplus
extern "C" {
#endif

typedef struct LsConv2d {
    int32_t N, H, W, Cin;  /* input x (N,H,W,Cin); Cin % 4 == 0                                                   */
    int32_t Cout, R, S;    /* filter; Cout % 4 == 0                                                               */
    int32_t stride, pad;   /* same in both directions; Conv2d: stride 1..8.  transposed: stride == R == S, pad 0  */
    int32_t transposed;    /* 0: nn.Conv2d   1: nn.ConvTranspose2d (non-overlapping: kernel == stride)            */
} LsConv2d;

/* output spatial size of the layer (host helper) */
LS_API int ls_conv2d_out_size(const LsConv2d* c, int32_t* out_h, int32_t* out_w);

/* y = act(conv(x, w) + bias).  bias may be NULL.  If y_pre != NULL the pre-activation values are stored there as well
 * (needed by the backward of GELU / SiLU); act is one of LS_ACT_* (ls_gemm.h).                                    */
LS_API int ls_conv2d_forward(const LsConv2d* c, const float* x, const float* w, const float* bias, float* y, float* y_pre,
                             int32_t act, void* stream /* cudaStream_t */);

/* dx = conv_transpose(dy, w): gradient w.r.t. the input (x-shaped).  Every element of dx is written.              */
LS_API int ls_conv2d_dgrad(const LsConv2d* c, const float* dy, const float* w, float* dx, void* stream);

/* dw = gradient w.r.t. the weight (weight-shaped, same physical layout as w); zero-filled here, split over the
 * pixel dimension and combined with red.global.add.                                                              */
LS_API int ls_conv2d_wgrad(const LsConv2d* c, const float* dy, const float* x, float* dw, void* stream);

/* ---- nearest-2x up-sampling followed by a 3x3 / stride 1 / pad 1 convolution -------------------------------------
 * diffusers' Upsample2D, the three up-samplers of the VAE decoder the reference wraps (autoencoder_kl.py:119-122 ->
 * Decoder.up_blocks[i].upsamplers[0]: F.interpolate(scale_factor=2, mode="nearest") then Conv2d(C, C, 3, padding=1)).
 * Per output parity class the 3x3 filter on the up-sampled image collapses to 2x2 taps with summed weights on the LOW
 * resolution input: 16 instead of 36 tap-MACs per input pixel and no up-sampled tensor in HBM.  `c` describes the
 * low-resolution input (N, H, W, Cin) and the 3x3 filter; y / dy are (N, 2H, 2W, Cout).
 * wk: caller-owned workspace of ls_upconv2x_workspace() floats, written by forward (folded weights) and read by dgrad;
 * scratch (wgrad): 16*Cout*Cin floats.                                                                               */
LS_API int ls_upconv2x_workspace(const LsConv2d* c, int64_t* floats);
LS_API int ls_upconv2x_forward(const LsConv2d* c, const float* x, const float* w, const float* bias, float* y, float* wk,
                               void* stream);
LS_API int ls_upconv2x_dgrad(const LsConv2d* c, const float* dy, const float* wk, float* dx, void* stream);
LS_API int ls_upconv2x_wgrad(const LsConv2d* c, const float* dy, const float* x, float* dw, float* scratch, void* stream);

/* ---- small NHWC helpers that ride along with the convolutions ------------------------------------------------- */
/* dx = dy * act'(pre)  (elementwise; the backward of a fused epilogue activation)                                */
LS_API int ls_act_backward(const float* dy, const float* pre, float* dx, int64_t n, int32_t act, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* LS_CONV_H */
