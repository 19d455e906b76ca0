/*
 * ls_raster.h -- C ABI of libls_raster.so, the sm_100a Gaussian rasterizer that
 * sits under the `diff_gaussian_rasterization` Python module.
 *
 * What each entry point replaces in the reference (Chrixtar/latentsplat):
 *   ls_raster_forward   <- the forward of `GaussianRasterizer(settings)(...)`,
 *                          /root/reference/src/model/decoder/cuda_splatting.py:146-158
 *                          (one call per view there; `n_views` views per call here, so the
 *                          Python loop at cuda_splatting.py:124-162 collapses to one launch
 *                          sequence).  Settings fields: cuda_splatting.py:132-145.
 *   ls_raster_backward  <- the autograd backward of the same call (gradients w.r.t.
 *                          means3D, means2D, shs | colors_precomp, features, opacities,
 *                          cov3D_precomp), triggered by ModelWrapper.manual_backward,
 *                          /root/reference/src/model/model_wrapper.py:440.
 *   ls_raster_sizes     <- the geometry/binning/image buffer sizing the reference's
 *                          extension does internally with torch resize lambdas [EXT].
 * The reference binds its rasterizer through a torch C++ extension (pybind11); the
 * binding a maintainer would write against THIS library is the ctypes stub in
 * INTEGRATION.md (no torch types cross this boundary).
 *
 * Conventions
 *   - All pointers are DEVICE pointers unless marked host.  The caller allocates and
 *     owns every buffer; the library never allocates or frees device memory and keeps
 *     no global state besides a thread-local error string.
 *   - All work is enqueued on `stream`; no call synchronises the device.
 *   - Return value 0 = ok, negative = error (message via ls_last_error()).  Nothing
 *     throws across the ABI.
 *   - fp32 everywhere; uint64 sort keys; int32 radii; uint32 counters.
 *   - Tiles are 16x16 pixels; T = ceil(W/16)*ceil(H/16) tiles per view.
 *   - Views are grouped into scenes: view v renders scene v / views_per_scene.  Per-Gaussian
 *     inputs are (S,G,...) with S = n_views / views_per_scene, so one scene rendered from
 *     several target cameras is NOT replicated per view (the reference materialises that
 *     repeat at decoder_splatting_cuda.py:71-86).  Input gradients are (S,G,...), summed
 *     over the views of a scene.
 */
#ifndef LS_RASTER_H
#define LS_RASTER_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LS_RASTER_ABI_VERSION 2
#if defined(__GNUC__)
#define LS_API __attribute__((visibility("default")))
#else
#define LS_API
#endif
#define LS_TILE 16
#define LS_GEOM_STRIDE 8          /* floats per Gaussian geometry record                 */
#define LS_MAX_VALUE_CHANNELS 16  /* colour (0|3) + feature channels blended in one pass */

enum { LS_COLOR_NONE = 0, LS_COLOR_PRECOMP = 1,
       LS_COLOR_SH = 2,       /* colour SH in the basis of the in-tree src/misc/sh_utils.py:42-97 (b1..3 = -C1 x, +C1 y, -C1 z) */
       LS_COLOR_SH_3DGS = 3   /* [EXT] colour SH in the coefficient order of graphdeco-inria/diff-gaussian-rasterization
                                 (b1..3 = -C1 y, +C1 z, -C1 x): the in-tree polynomials at (y, z, x), except k = 14 where the
                                 in-tree file has z(zz - xx) and 3DGS z(xx - yy); degree 4 is an extension.  The reference
                                 hands `shs` to its fork's CUDA (cuda_splatting.py:91,146), whose basis cannot be inspected
                                 here; LS_SH_BASIS=intree|3dgs selects between the two candidates in the Python layer.   */
};
enum { LS_FEATURE_NONE = 0, LS_FEATURE_PRECOMP = 1, LS_FEATURE_SH = 2 };
/* forward stages (bit mask); each may be launched separately so that a profiler can bracket it */
enum { LS_STAGE_GEOMETRY = 1,  /* preprocess + per-tile count + exclusive scan (writes stats[0..2]) */
       LS_STAGE_SCATTER = 2,   /* (depth|id) keys into the tile segments                            */
       LS_STAGE_SORT = 4,      /* per-tile radix sort                                               */
       LS_STAGE_BLEND = 8,     /* front-to-back compositing                                         */
       LS_STAGE_RENDER = 2 | 4 | 8,
       LS_STAGE_ALL = 15 };
/* backward stages */
enum { LS_BWD_BLEND = 1,       /* back-to-front replay -> per-Gaussian gradient records             */
       LS_BWD_GEOMETRY = 2,    /* records -> input gradients                                        */
       LS_BWD_ALL = 3 };

typedef struct LsRasterScene {
    int32_t n_views;           /* V                                                          */
    int32_t views_per_scene;   /* V % views_per_scene == 0; S = V / views_per_scene          */
    int32_t G, H, W;
    int32_t C;                 /* feature channels (0 = none)                                */
    int32_t color_mode;        /* LS_COLOR_*                                                 */
    int32_t sh_degree;         /* colour SH degree 0..4 (cuda_splatting.py:141)              */
    int32_t feature_mode;      /* LS_FEATURE_*; _SH evaluates 0.5+eval_sh in-kernel, fusing
                                  cuda_splatting.py:94-101 (sh_utils.py:42-97)               */
    int32_t feature_sh_degree; /* 0..4                                                       */
    const float* means3D;      /* (S,G,3)                                                    */
    const float* cov3D;        /* (S,G,6) upper triangle 00 01 02 11 12 22 (cuda_splatting.py:148) */
    const float* opacity;      /* (S,G)                                                      */
    const float* color;        /* PRECOMP: (S,G,3); SH: (S,G,(deg+1)^2,3) (cuda_splatting.py:91) */
    const float* feature;      /* PRECOMP: (S,G,C); SH: (S,G,C,(fdeg+1)^2)                   */
    const float* viewmatrix;   /* (V,16) transposed world->view (cuda_splatting.py:116)      */
    const float* projmatrix;   /* (V,16) transposed full projection (:117)                   */
    const float* campos;       /* (V,3)                                                      */
    const float* tanfov;       /* (V,2) tan(fov_x/2), tan(fov_y/2) -- device, no .item()     */
    const float* bg;           /* (V,3) colour background (features composite over 0)        */
    const float* scene_scale;  /* (V) or NULL: means*=s, cov*=s*s (cuda_splatting.py:75-82)  */
} LsRasterScene;

/* State written by forward and read by backward. */
typedef struct LsRasterState {
    float*    geom;          /* (V,G,8): x, y, A, B | Cc, opacity, depth, ext  with ext = two fp16
                                half-extents of the alpha >= 1/255 region (conservative), and
                                (A,B,Cc) = (-0.5 cxx, -cxy, -0.5 cyy) * log2(e), the conic
                                pre-scaled so that alpha = opacity * exp2(A dx^2 + B dx dy + Cc dy^2) */
    float*    chan;          /* (V,G,chan_stride): colour(0|3) then C features               */
    int32_t*  radii;         /* (V,G)  0 = culled                                            */
    uint32_t* tiles_touched; /* (V,G)                                                        */
    uint8_t*  clamped;       /* (V,G) bit c set when colour channel c was clamped at 0       */
    uint32_t* tile_count;    /* (V*T) scratch                                                */
    uint32_t* tile_offsets;  /* (V*T+1) exclusive prefix; [V*T] = num_rendered               */
    uint32_t* stats;         /* (4): num_rendered, max tile count, overflow flag, 0          */
    uint64_t* keys;          /* (capacity) depth_bits<<32 | gaussian id, sorted per tile     */
    uint64_t* keys_tmp;      /* (capacity) scratch for tiles longer than the smem tier       */
    int64_t   capacity;      /* entries in keys / keys_tmp                                   */
    float*    final_T;       /* (V,H,W)                                                      */
    uint32_t* n_contrib;     /* (V,H,W)                                                      */
    int32_t   chan_stride;   /* floats per chan record = round_up(max(1, ncolor + C), 4)     */
    int32_t   sort_smem_keys;/* keys of one tile held in shared memory by the per-tile sort
                                (0 = default 4096); longer tiles sort through keys_tmp       */
    /* The per-tile Gaussian QUEUES in list order, written by the sort stage (a permute-after-sort): the blend
     * kernels stage them with 1-D bulk TMA copies (cp.async.bulk + mbarrier) instead of per-thread gathers.    */
    float*    sorted_cull;   /* (capacity, 4): x, y, fp16x2 half-extents, Gaussian id (bit pattern) -- what the
                                per-warp compaction pass reads, 16 B per entry, bank-conflict free              */
    float*    sorted_rec;    /* (capacity, rec_stride): the 8-float geometry record followed by the chan record */
    int32_t   rec_stride;    /* 8 + chan_stride                                                                 */
    int32_t   reserved1;
} LsRasterState;

typedef struct LsRasterImages {
    float* color;   /* (V,3,H,W) or NULL when color_mode == NONE */
    float* feature; /* (V,C,H,W) or NULL when C == 0             */
    float* alpha;   /* (V,H,W) accumulated alpha ("mask")        */
    float* depth;   /* (V,H,W) sum_i depth_i alpha_i T_i         */
} LsRasterImages;

typedef struct LsRasterGrads {
    /* upstream gradients; any may be NULL (= zero) */
    const float* dL_dcolor;   /* (V,3,H,W) */
    const float* dL_dfeature; /* (V,C,H,W) */
    const float* dL_dalpha;   /* (V,H,W)   */
    const float* dL_ddepth;   /* (V,H,W)   */
    /* scratch */
    float* dL_drecord;        /* (V,G,grad_stride) zero-filled by the library:
                                 mean2D.x, .y, conic xx, xy, yy, opacity, depth, chan...     */
    int32_t grad_stride;      /* round_up(7 + ncolor + C, 4)                                 */
    int32_t reserved0;
    /* outputs, zero-filled by the library, summed over the views of each scene */
    float* dL_dmeans3D;       /* (S,G,3) */
    float* dL_dcov3D;         /* (S,G,6) */
    float* dL_dopacity;       /* (S,G)   */
    float* dL_dcolor_in;      /* PRECOMP: (S,G,3); SH: (S,G,(deg+1)^2,3); NULL if none       */
    float* dL_dfeature_in;    /* PRECOMP: (S,G,C); SH: (S,G,C,(fdeg+1)^2); NULL if none      */
    float* dL_dmeans2D;       /* (V,G,3) screen-space sink (z = 0), per view; may be NULL    */
    /* Row pitch (floats) of dL_dcolor_in / dL_dfeature_in, 0 = dense.  The SH coefficient rows are 75 / 36 floats (300 / 144
     * B): written densely, their 60-B chunks straddle 32-B sectors and L2 read-modify-writes them (ncu r01: 2.1x the
     * algorithmic DRAM traffic).  A pitch that is a multiple of 8 floats makes every chunk whole sectors; the caller hands the
     * [:, :row] view of the padded buffer to autograd.                                                                        */
    int32_t color_grad_pitch, feature_grad_pitch;
} LsRasterGrads;

/* Buffer sizes (in elements) for a given problem; `out` is a host pointer. */
typedef struct LsRasterSizes {
    int64_t n_scenes, tiles_per_view, geom, chan, per_view_gaussian, tile_slots, pixels, grad_record;
    int32_t chan_stride, grad_stride, n_color, n_value_channels;
    int32_t rec_stride, reserved0;
} LsRasterSizes;

LS_API int ls_raster_sizes(const LsRasterScene* scene, LsRasterSizes* out /* host */);

/* stages: LS_STAGE_GEOMETRY writes stats[0..2]; the LS_STAGE_RENDER stages need
 * state->capacity >= stats[0]; pass LS_STAGE_ALL for a single sync-free call with a
 * caller-chosen capacity (stats[2] is set to 1 when num_rendered exceeds it; entries past
 * the capacity are dropped). */
LS_API int ls_raster_forward(const LsRasterScene* scene, const LsRasterState* state, const LsRasterImages* images,
                      int32_t stages, void* stream /* cudaStream_t */);

LS_API int ls_raster_backward(const LsRasterScene* scene, const LsRasterState* state, const LsRasterGrads* grads,
                       int32_t stages, void* stream /* cudaStream_t */);

LS_API const char* ls_last_error(void); /* thread-local, never NULL */
LS_API int ls_raster_abi_version(void);

/* 1 when the specialised coefficient backward (colour SH degree 4 in the in-tree basis, 4 or 8 feature channels of SH degree 2,
 * G % 4 == 0) applies to this scene: the caller should then allocate DENSE SH gradient rows (color_grad_pitch = feature_grad_pitch
 * = 0) -- the warp's 32 rows leave shared memory as one bulk copy; otherwise rows padded to 8 floats are the faster layout. */
LS_API int ls_raster_dense_sh_grads(const LsRasterScene* scene);

#ifdef __cplusplus
}
#endif
#endif /* LS_RASTER_H */
