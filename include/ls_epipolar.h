/*
 * ls_epipolar.h -- C ABI of the fused epipolar-line feature gather (+ depth encoding) of the encoder.
 *
 * What it replaces in the reference (Chrixtar/latentsplat):
 *   src/model/encoder/epipolar/epipolar_sampler.py:96-112     transpose -> F.grid_sample(bilinear, zeros padding,
 *       align_corners=False) -> rearrange -> transpose -> multiply by projection["overlaps_image"]
 *   src/model/encoder/epipolar/epipolar_transformer.py:121-122   q = sampling.features + depth_encoding(depths[..., None])
 *       (depth_encoding = PositionalEncoding(num_octaves) -> Linear, :51-54; positional_encoding.py:8-36)
 *
 *   z[row, s, :] = valid[row] * bilinear(feat[image[row]], xy[row, s])  +  (We . pe(depth[row, s]) + be)
 *
 * A "row" is one (scene, view, other view, ray) epipolar line with `samples` points on it; `image[row]` is the flattened
 * (scene * views + other view) index of the feature map the line lies in, so the reference's two index "transposes"
 * disappear.  Features are CHANNELS-LAST here: feat (images, height, width, 128).  xy is in normalised image coordinates
 * ([0,1]^2, x right, y down) exactly as EpipolarSampling.xy_sample.  Sample positions and depths are camera geometry; no
 * gradient is produced for them.
 *
 * Conventions as ls_raster.h: device pointers, caller-owned buffers, work enqueued on `stream`, no sync, 0 / negative
 * return + ls_last_error().  feat / z / dz / dfeat must be 16-byte aligned.  backward ACCUMULATES into dfeat, dWe, dbe
 * (the caller zero-fills them).
 */
#ifndef LS_EPIPOLAR_H
#define LS_EPIPOLAR_H

#include <stdint.h>

#include "ls_raster.h" /* LS_API, ls_last_error */

#ifdef __cplusplus
extern "C" {
#endif

typedef struct LsEpipolarGather {
    int32_t rows;            /* epipolar lines: scenes * views * other_views * rays                     */
    int32_t samples;         /* points per line, 1..32 (epipolar_transformer num_samples)                */
    int32_t images;          /* feature maps: scenes * views                                             */
    int32_t height, width;   /* feature-map size (after the transformer's downscaler)                    */
    int32_t channels;        /* must be 128                                                              */
    int32_t encoding_width;  /* 0: no depth encoding;  20: 10 octaves x (sin, cos), the shipped config   */
    const float* xy;         /* (rows, samples, 2)                                                       */
    const float* depth;      /* (rows, samples) relative disparity in [0,1]; NULL iff encoding_width==0  */
    const int32_t* image;    /* (rows)                                                                   */
    const float* valid;      /* (rows) 1.0 / 0.0 = projection["overlaps_image"]                          */
} LsEpipolarGather;

/* feat (images, height, width, 128); We (128, encoding_width) row-major = depth_encoding[1].weight; be (128);
 * z (rows, samples, 128). */
LS_API int ls_epipolar_gather_forward(const LsEpipolarGather* args, const float* feat, const float* We, const float* be,
                                      float* z, void* stream /* cudaStream_t */);
/* dz (rows, samples, 128) -> dfeat (images, height, width, 128) += , dWe (128, encoding_width) +=, dbe (128) += */
LS_API int ls_epipolar_gather_backward(const LsEpipolarGather* args, const float* dz, float* dfeat, float* dWe, float* dbe,
                                       void* stream /* cudaStream_t */);

#ifdef __cplusplus
}
#endif
#endif /* LS_EPIPOLAR_H */
