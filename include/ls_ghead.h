/*
 * ls_ghead.h -- C ABI of the fused per-ray tail of the epipolar encoder: depth-bucket sampling + Gaussian adapter.
 *
 * What it replaces in the reference (Chrixtar/latentsplat), executed there as ~60 (forward) + ~100 (backward) eager
 * torch kernels over (rays x samples) tensors:
 *   DepthPredictorMonocular.forward   src/model/encoder/epipolar/depth_predictor_monocular.py:37-81
 *   sample_discrete_distribution /    src/misc/discrete_probability_distribution.py:7-33
 *   gather_discrete_topk
 *   relative_disparity_to_depth       src/model/encoder/epipolar/conversions.py:5-14
 *   xy offsets, map_pdf_to_opacity    src/model/encoder/encoder_epipolar.py:113-126, 183-190
 *   GaussianAdapter.forward           src/model/encoder/common/gaussian_adapter.py:63-114 (+ gaussians.py:8-44,
 *                                     get_world_rays src/geometry/projection.py:98-121)
 * One warp per ray (= context-view pixel), lane l = depth bucket l.  Inputs are the outputs of the two Linear heads
 * (depth logits, raw Gaussian parameters), the cameras and the uniform random numbers of the bucket draw; outputs are
 * the world-space Gaussians of the ray's `samples` depth samples, laid out exactly as VariationalGaussians.flatten
 * orders them ((view, ray, sample) with sample fastest).
 *
 * Conventions as ls_raster.h: device pointers, caller-owned buffers, work enqueued on `stream`, no sync, 0 / negative
 * return + ls_last_error().  fp32 everywhere; `index` int32.
 */
#ifndef LS_GHEAD_H
#define LS_GHEAD_H

#include <stdint.h>

#include "ls_raster.h" /* LS_API, ls_last_error */

#ifdef __cplusplus
extern "C" {
#endif

typedef struct LsGaussianHead {
    int64_t rays;            /* B * V * rays_per_view                                                           */
    int32_t rays_per_view;   /* height * width                                                                  */
    int32_t width, height;   /* pixel grid of a context view (xy = pixel centre + sigmoid offset)               */
    int32_t samples;         /* Gaussians per ray (gaussians_per_pixel, 1..4); 1 in deterministic mode          */
    int32_t buckets;         /* depth buckets (num_monocular_samples); must be 32                               */
    int32_t d_color;         /* floats of the colour SH block of a raw row (3 * 25)                             */
    int32_t d_feature;       /* floats of the feature SH block (channels * 9)                                   */
    int32_t deterministic;   /* 0: draw buckets with u (searchsorted on the cdf); 1: top-1 bucket               */
    float scale_min, scale_max;     /* gaussian_adapter.py: scale = min + (max - min) sigmoid(raw)              */
    float opacity_exponent;  /* 2^x of map_pdf_to_opacity (1 = identity)                                        */
    float inv_gpp;           /* 1 / cfg.gaussians_per_pixel                                                     */
    const float* dlog;       /* (rays, 2 * buckets): depth head output, (bucket, {pdf logit, offset logit})     */
    const float* raw;        /* (rays, 9 + d_color + d_feature): ox oy | s0 s1 s2 | qx qy qz qw | SH blocks     */
    const float* u;          /* (rays, samples) uniform [0,1) draws, or NULL when deterministic                 */
    const float* extrinsics; /* (views, 4, 4) camera-to-world                                                   */
    const float* intrinsics; /* (views, 3, 3) normalised                                                        */
    const float* near;       /* (views)                                                                         */
    const float* far;        /* (views)                                                                         */
} LsGaussianHead;

typedef struct LsGaussianHeadOut {      /* G = rays * samples Gaussians */
    float* means;            /* (G, 3)          */
    float* covariances;      /* (G, 3, 3)       */
    float* opacity;          /* (G)             */
    float* color_sh;         /* (G, d_color)    */
    float* feature_sh;       /* (G, d_feature)  */
    int32_t* index;          /* (G) sampled bucket, kept for the backward pass */
} LsGaussianHeadOut;

typedef struct LsGaussianHeadGrad {
    const int32_t* index;        /* (G) from forward */
    const float* d_means;        /* (G, 3)           */
    const float* d_covariances;  /* (G, 3, 3)        */
    const float* d_opacity;      /* (G)              */
    const float* d_color_sh;     /* (G, d_color)     */
    const float* d_feature_sh;   /* (G, d_feature)   */
    float* d_dlog;               /* (rays, 2 * buckets)            written */
    float* d_raw;                /* (rays, 9 + d_color + d_feature) written */
} LsGaussianHeadGrad;

LS_API int ls_gaussian_head_forward(const LsGaussianHead* args, const LsGaussianHeadOut* out, void* stream /* cudaStream_t */);
LS_API int ls_gaussian_head_backward(const LsGaussianHead* args, const LsGaussianHeadGrad* grads, void* stream /* cudaStream_t */);

/* Reparameterised sample of the variational Gaussians' feature harmonics (VariationalGaussians.sample ->
 * DiagonalGaussianDistribution.sample, /root/reference/src/model/diagonal_gaussian_distribution.py:30-36, 78-81; called at
 * model_wrapper.py:362): params (rows, 2*half) = [mean | logvar] per row, eps (rows, half) standard-normal draws (torch's RNG),
 *   out = mean + exp(0.5 * clamp(logvar, lo, hi)) * eps          -- one pass instead of clamp / exp / exp / mul / add
 *   d_params = [g | g * eps * 0.5 * std * 1(lo <= logvar <= hi)]  -- written as one (rows, 2*half) tensor
 * half % 4 == 0, 16-byte aligned pointers. */
LS_API int ls_reparam_forward(const float* params, const float* eps, float* out, int64_t rows, int32_t half, float lo, float hi,
                              void* stream /* cudaStream_t */);
LS_API int ls_reparam_backward(const float* params, const float* eps, const float* g, float* d_params, int64_t rows, int32_t half,
                               float lo, float hi, void* stream /* cudaStream_t */);

#ifdef __cplusplus
}
#endif
#endif /* LS_GHEAD_H */
