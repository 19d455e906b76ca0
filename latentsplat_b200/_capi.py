"""ctypes binding of libls_raster.so (include/ls_raster.h).  No torch types cross the ABI:
tensors are handed over as raw device pointers + sizes, the stream as a cudaStream_t.

There is NO CPU fallback: if the library is missing or a call fails, this raises.
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path

from . import _build

_fp = C.POINTER(C.c_float)
_u32p = C.POINTER(C.c_uint32)


class LsRasterScene(C.Structure):
    _fields_ = [
        ("n_views", C.c_int32), ("views_per_scene", C.c_int32), ("G", C.c_int32), ("H", C.c_int32),
        ("W", C.c_int32), ("C", C.c_int32), ("color_mode", C.c_int32), ("sh_degree", C.c_int32),
        ("feature_mode", C.c_int32), ("feature_sh_degree", C.c_int32),
        ("means3D", C.c_void_p), ("cov3D", C.c_void_p), ("opacity", C.c_void_p), ("color", C.c_void_p),
        ("feature", C.c_void_p), ("viewmatrix", C.c_void_p), ("projmatrix", C.c_void_p), ("campos", C.c_void_p),
        ("tanfov", C.c_void_p), ("bg", C.c_void_p), ("scene_scale", C.c_void_p),
    ]


class LsRasterState(C.Structure):
    _fields_ = [
        ("geom", C.c_void_p), ("chan", C.c_void_p), ("radii", C.c_void_p), ("tiles_touched", C.c_void_p),
        ("clamped", C.c_void_p), ("tile_count", C.c_void_p), ("tile_offsets", C.c_void_p), ("stats", C.c_void_p),
        ("keys", C.c_void_p), ("keys_tmp", C.c_void_p), ("capacity", C.c_int64), ("final_T", C.c_void_p),
        ("n_contrib", C.c_void_p), ("chan_stride", C.c_int32), ("sort_smem_keys", C.c_int32),
        ("sorted_cull", C.c_void_p), ("sorted_rec", C.c_void_p), ("rec_stride", C.c_int32), ("reserved1", C.c_int32),
    ]


class LsRasterImages(C.Structure):
    _fields_ = [("color", C.c_void_p), ("feature", C.c_void_p), ("alpha", C.c_void_p), ("depth", C.c_void_p)]


class LsRasterGrads(C.Structure):
    _fields_ = [
        ("dL_dcolor", C.c_void_p), ("dL_dfeature", C.c_void_p), ("dL_dalpha", C.c_void_p), ("dL_ddepth", C.c_void_p),
        ("dL_drecord", C.c_void_p), ("grad_stride", C.c_int32), ("reserved0", C.c_int32),
        ("dL_dmeans3D", C.c_void_p), ("dL_dcov3D", C.c_void_p), ("dL_dopacity", C.c_void_p),
        ("dL_dcolor_in", C.c_void_p), ("dL_dfeature_in", C.c_void_p), ("dL_dmeans2D", C.c_void_p),
        ("color_grad_pitch", C.c_int32), ("feature_grad_pitch", C.c_int32),
    ]


class LsRasterSizes(C.Structure):
    _fields_ = [
        ("n_scenes", C.c_int64), ("tiles_per_view", C.c_int64), ("geom", C.c_int64), ("chan", C.c_int64),
        ("per_view_gaussian", C.c_int64), ("tile_slots", C.c_int64), ("pixels", C.c_int64),
        ("grad_record", C.c_int64), ("chan_stride", C.c_int32), ("grad_stride", C.c_int32),
        ("n_color", C.c_int32), ("n_value_channels", C.c_int32), ("rec_stride", C.c_int32), ("reserved0", C.c_int32),
    ]


class LsGemmArgs(C.Structure):
    _fields_ = [
        ("M", C.c_int32), ("N", C.c_int32), ("K", C.c_int32), ("a_mn_major", C.c_int32), ("b_mn_major", C.c_int32),
        ("act", C.c_int32), ("split_k", C.c_int32), ("accumulate", C.c_int32),
        ("lda", C.c_int64), ("ldb", C.c_int64), ("ldc", C.c_int64),
        ("A", C.c_void_p), ("B", C.c_void_p), ("C", C.c_void_p), ("bias", C.c_void_p),
        ("residual", C.c_void_p), ("ldr", C.c_int64), ("pre_out", C.c_void_p),
    ]


class LsEpipolarGather(C.Structure):     # include/ls_epipolar.h
    _fields_ = [(n, C.c_int32) for n in ("rows", "samples", "images", "height", "width", "channels", "encoding_width")] + \
               [(n, C.c_void_p) for n in ("xy", "depth", "image", "valid")]


class LsGroupNorm(C.Structure):          # include/ls_norm.h
    _fields_ = [("N", C.c_int32), ("C", C.c_int32), ("G", C.c_int32), ("act", C.c_int32), ("HW", C.c_int64), ("eps", C.c_float),
                ("x", C.c_void_p), ("gamma", C.c_void_p), ("beta", C.c_void_p), ("stats", C.c_void_p)]


class LsConv2d(C.Structure):             # include/ls_conv.h
    _fields_ = [(n, C.c_int32) for n in ("N", "H", "W", "Cin", "Cout", "R", "S", "stride", "pad", "transposed")]


class LsGaussianHead(C.Structure):      # include/ls_ghead.h
    _fields_ = [("rays", C.c_int64)] + \
               [(n, C.c_int32) for n in ("rays_per_view", "width", "height", "samples", "buckets", "d_color", "d_feature",
                                         "deterministic")] + \
               [(n, C.c_float) for n in ("scale_min", "scale_max", "opacity_exponent", "inv_gpp")] + \
               [(n, C.c_void_p) for n in ("dlog", "raw", "u", "extrinsics", "intrinsics", "near", "far")]


class LsGaussianHeadOut(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("means", "covariances", "opacity", "color_sh", "feature_sh", "index")]


class LsGaussianHeadGrad(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("index", "d_means", "d_covariances", "d_opacity", "d_color_sh", "d_feature_sh",
                                          "d_dlog", "d_raw")]


class LsFmha(C.Structure):              # include/ls_fmha.h
    _fields_ = [("B", C.c_int32), ("H", C.c_int32), ("L", C.c_int32), ("D", C.c_int32), ("scale", C.c_float), ("reserved0", C.c_int32),
                ("ld_q", C.c_int64), ("ld_k", C.c_int64), ("ld_v", C.c_int64), ("ld_o", C.c_int64),
                ("q", C.c_void_p), ("k", C.c_void_p), ("v", C.c_void_p), ("o", C.c_void_p), ("lse", C.c_void_p)]


ACT_NONE, ACT_RELU, ACT_GELU, ACT_SILU, ACT_LRELU = 0, 1, 2, 3, 4
MAX_VALUE_CHANNELS = 16          # LS_MAX_VALUE_CHANNELS (include/ls_raster.h): colour + feature channels blended in one pass
COLOR_NONE, COLOR_PRECOMP, COLOR_SH, COLOR_SH_3DGS = 0, 1, 2, 3
FEATURE_NONE, FEATURE_PRECOMP, FEATURE_SH = 0, 1, 2
STAGE_GEOMETRY, STAGE_SCATTER, STAGE_SORT, STAGE_BLEND, STAGE_RENDER, STAGE_ALL = 1, 2, 4, 8, 14, 15
BWD_BLEND, BWD_GEOMETRY, BWD_ALL = 1, 2, 3
ABI_VERSION = 2
EXPORTS = ("ls_raster_sizes", "ls_raster_forward", "ls_raster_backward", "ls_last_error", "ls_raster_abi_version", "ls_raster_dense_sh_grads",
           "ls_gemm_tf32", "ls_sq_attention_forward", "ls_sq_attention_backward",
           "ls_absorbed_attention_forward", "ls_absorbed_attention_backward",
           "ls_epipolar_gather_forward", "ls_epipolar_gather_backward", "ls_groupnorm_forward", "ls_groupnorm_backward",
           "ls_layernorm_forward", "ls_layernorm_backward", "ls_conv_bias_add", "ls_conv_bias_grad", "ls_col_sum",
           "ls_groupnorm_nhwc_forward", "ls_groupnorm_nhwc_backward",
           "ls_conv2d_out_size", "ls_conv2d_forward", "ls_conv2d_dgrad", "ls_conv2d_wgrad", "ls_act_backward",
           "ls_upconv2x_workspace", "ls_upconv2x_forward", "ls_upconv2x_dgrad", "ls_upconv2x_wgrad",
           "ls_gaussian_head_forward", "ls_gaussian_head_backward", "ls_reparam_forward", "ls_reparam_backward", "ls_fmha_forward", "ls_fmha_backward",
           "ls_softmax_rows_forward", "ls_softmax_rows_backward")

_lib = None
# algorithmic FLOPs handed to our tensor-core kernels, per family (bench.py: in-step aggregate = d FLOPS / d kernel time)
FLOPS = {"gemm": 0.0, "conv": 0.0, "fmha": 0.0}
KERNEL_LAUNCHES = [0]   # running count of OUR kernel launches (bench.py reports the per-step delta as gpu_launches)


def lib_path() -> Path:
    return _build.LIB


def load() -> C.CDLL:
    """Load libls_raster.so (built in-tree by latentsplat_b200._build).  Raises if absent."""
    global _lib
    if _lib is not None:
        return _lib
    path = lib_path()
    if not path.exists():
        raise RuntimeError(
            f"{path} is missing: build it with `python -m latentsplat_b200._build` "
            "(or __graft_entry__.build()).  latentsplat_b200 has no CPU fallback.")
    lib = C.CDLL(str(path))
    lib.ls_last_error.restype = C.c_char_p
    lib.ls_raster_abi_version.restype = C.c_int
    lib.ls_raster_sizes.restype = C.c_int
    lib.ls_raster_sizes.argtypes = [C.POINTER(LsRasterScene), C.POINTER(LsRasterSizes)]
    lib.ls_raster_forward.restype = C.c_int
    lib.ls_raster_forward.argtypes = [C.POINTER(LsRasterScene), C.POINTER(LsRasterState), C.POINTER(LsRasterImages),
                                      C.c_int32, C.c_void_p]
    lib.ls_raster_backward.restype = C.c_int
    lib.ls_raster_backward.argtypes = [C.POINTER(LsRasterScene), C.POINTER(LsRasterState), C.POINTER(LsRasterGrads),
                                       C.c_int32, C.c_void_p]
    lib.ls_gemm_tf32.restype = C.c_int
    lib.ls_gemm_tf32.argtypes = [C.POINTER(LsGemmArgs), C.c_void_p]
    lib.ls_sq_attention_forward.restype = C.c_int
    lib.ls_sq_attention_forward.argtypes = [C.c_void_p] * 4 + [C.c_int32] * 4 + [C.c_float, C.c_void_p]
    lib.ls_sq_attention_backward.restype = C.c_int
    lib.ls_sq_attention_backward.argtypes = [C.c_void_p] * 6 + [C.c_int32] * 4 + [C.c_float, C.c_void_p]
    lib.ls_absorbed_attention_forward.restype = C.c_int
    lib.ls_absorbed_attention_forward.argtypes = [C.c_void_p] * 4 + [C.c_int32] * 4 + [C.c_float, C.c_void_p]
    lib.ls_absorbed_attention_backward.restype = C.c_int
    lib.ls_absorbed_attention_backward.argtypes = [C.c_void_p] * 6 + [C.c_int32] * 4 + [C.c_float, C.c_void_p]
    lib.ls_epipolar_gather_forward.restype = C.c_int
    lib.ls_epipolar_gather_forward.argtypes = [C.POINTER(LsEpipolarGather)] + [C.c_void_p] * 5
    lib.ls_epipolar_gather_backward.restype = C.c_int
    lib.ls_epipolar_gather_backward.argtypes = [C.POINTER(LsEpipolarGather)] + [C.c_void_p] * 5
    lib.ls_groupnorm_forward.restype = C.c_int
    lib.ls_groupnorm_forward.argtypes = [C.POINTER(LsGroupNorm), C.c_void_p, C.c_void_p]
    lib.ls_groupnorm_backward.restype = C.c_int
    lib.ls_groupnorm_backward.argtypes = [C.POINTER(LsGroupNorm)] + [C.c_void_p] * 4
    lib.ls_groupnorm_nhwc_forward.restype = C.c_int
    lib.ls_groupnorm_nhwc_forward.argtypes = [C.POINTER(LsGroupNorm), C.c_void_p, C.c_void_p]
    lib.ls_groupnorm_nhwc_backward.restype = C.c_int
    lib.ls_groupnorm_nhwc_backward.argtypes = [C.POINTER(LsGroupNorm)] + [C.c_void_p] * 4
    lib.ls_layernorm_forward.restype = C.c_int
    lib.ls_layernorm_forward.argtypes = [C.c_void_p] * 5 + [C.c_int64, C.c_int32, C.c_float, C.c_void_p]
    lib.ls_layernorm_backward.restype = C.c_int
    lib.ls_layernorm_backward.argtypes = [C.c_void_p] * 7 + [C.c_int64, C.c_int32, C.c_void_p]
    for fn in (lib.ls_conv_bias_add, lib.ls_conv_bias_grad):
        fn.restype = C.c_int
        fn.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_int64, C.c_void_p]
    lib.ls_conv2d_out_size.restype = C.c_int
    lib.ls_conv2d_out_size.argtypes = [C.POINTER(LsConv2d), C.POINTER(C.c_int32), C.POINTER(C.c_int32)]
    lib.ls_conv2d_forward.restype = C.c_int
    lib.ls_conv2d_forward.argtypes = [C.POINTER(LsConv2d)] + [C.c_void_p] * 5 + [C.c_int32, C.c_void_p]
    lib.ls_conv2d_dgrad.restype = C.c_int
    lib.ls_conv2d_dgrad.argtypes = [C.POINTER(LsConv2d)] + [C.c_void_p] * 4
    lib.ls_conv2d_wgrad.restype = C.c_int
    lib.ls_conv2d_wgrad.argtypes = [C.POINTER(LsConv2d)] + [C.c_void_p] * 4
    lib.ls_upconv2x_workspace.restype = C.c_int
    lib.ls_upconv2x_workspace.argtypes = [C.POINTER(LsConv2d), C.POINTER(C.c_int64)]
    lib.ls_upconv2x_forward.restype = C.c_int
    lib.ls_upconv2x_forward.argtypes = [C.POINTER(LsConv2d)] + [C.c_void_p] * 6
    lib.ls_upconv2x_dgrad.restype = C.c_int
    lib.ls_upconv2x_dgrad.argtypes = [C.POINTER(LsConv2d)] + [C.c_void_p] * 4
    lib.ls_upconv2x_wgrad.restype = C.c_int
    lib.ls_upconv2x_wgrad.argtypes = [C.POINTER(LsConv2d)] + [C.c_void_p] * 5
    lib.ls_gaussian_head_forward.restype = C.c_int
    lib.ls_gaussian_head_forward.argtypes = [C.POINTER(LsGaussianHead), C.POINTER(LsGaussianHeadOut), C.c_void_p]
    lib.ls_gaussian_head_backward.restype = C.c_int
    lib.ls_gaussian_head_backward.argtypes = [C.POINTER(LsGaussianHead), C.POINTER(LsGaussianHeadGrad), C.c_void_p]
    lib.ls_reparam_forward.restype = C.c_int
    lib.ls_reparam_forward.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_float, C.c_float, C.c_void_p]
    lib.ls_reparam_backward.restype = C.c_int
    lib.ls_reparam_backward.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_float, C.c_float, C.c_void_p]
    lib.ls_raster_dense_sh_grads.restype = C.c_int
    lib.ls_raster_dense_sh_grads.argtypes = [C.POINTER(LsRasterScene)]
    lib.ls_fmha_forward.restype = C.c_int
    lib.ls_fmha_forward.argtypes = [C.POINTER(LsFmha), C.c_void_p]
    lib.ls_fmha_backward.restype = C.c_int
    lib.ls_fmha_backward.argtypes = [C.POINTER(LsFmha)] + [C.c_void_p] * 6
    lib.ls_softmax_rows_forward.restype = C.c_int
    lib.ls_softmax_rows_forward.argtypes = [C.c_void_p, C.c_int64, C.c_int32, C.c_int64, C.c_float, C.c_void_p]
    lib.ls_softmax_rows_backward.restype = C.c_int
    lib.ls_softmax_rows_backward.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_int64, C.c_float, C.c_void_p]
    lib.ls_act_backward.restype = C.c_int
    lib.ls_act_backward.argtypes = [C.c_void_p] * 3 + [C.c_int64, C.c_int32, C.c_void_p]
    lib.ls_col_sum.restype = C.c_int
    lib.ls_col_sum.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_int64, C.c_void_p]
    if lib.ls_raster_abi_version() != ABI_VERSION:
        raise RuntimeError(f"libls_raster.so ABI {lib.ls_raster_abi_version()} != binding {ABI_VERSION}; rebuild")
    _lib = lib
    return lib


def check(rc: int, what: str) -> None:
    if rc != 0:
        raise RuntimeError(f"{what} failed: {load().ls_last_error().decode()}")
