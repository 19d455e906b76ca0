"""Fused per-ray tail of the epipolar encoder (libls_raster.so::ls_gaussian_head_*, include/ls_ghead.h).

depth logits + raw Gaussian parameters + cameras (+ the uniform draws of the bucket sampling) -> world-space means,
covariances, opacities and the per-sample copies of the SH coefficient rows, one kernel forward and one backward, instead of
the reference's DepthPredictorMonocular -> offsets -> GaussianAdapter chain of eager ops
(/root/reference/src/model/encoder/epipolar/depth_predictor_monocular.py:37-81, encoder_epipolar.py:176-242,
common/gaussian_adapter.py:63-114).  CUDA fp32 only; the encoder keeps the explicit torch sequence for everything else.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import torch
from torch import Tensor

from . import _capi

ENABLED = True      # set False for A/B comparisons against the explicit torch sequence


def supported(dlog: Tensor, raw: Tensor, buckets: int, surfaces: int, samples: int, use_transmittance: bool,
              predict_opacity: bool) -> bool:
    return (ENABLED and dlog.is_cuda and dlog.dtype == raw.dtype == torch.float32 and buckets == 32 and surfaces == 1
            and 1 <= samples <= 4 and not use_transmittance and not predict_opacity and dlog.shape[-1] == 64)


class _GaussianHead(torch.autograd.Function):
    @staticmethod
    def forward(ctx, dlog: Tensor, raw: Tensor, u: Optional[Tensor], extrinsics: Tensor, intrinsics: Tensor, near: Tensor,
                far: Tensor, height: int, width: int, samples: int, d_color: int, d_feature: int, scale_min: float,
                scale_max: float, opacity_exponent: float, inv_gpp: float):
        b, v, r = dlog.shape[:3]
        rays = b * v * r
        dev = dlog.device
        dlog, raw = dlog.contiguous(), raw.contiguous()
        ext, intr = extrinsics.contiguous().float(), intrinsics.contiguous().float()
        near, far = near.contiguous().float(), far.contiguous().float()
        u = None if u is None else u.contiguous()
        G = rays * samples
        e = lambda *shape, dtype=torch.float32: torch.empty(shape, dtype=dtype, device=dev)
        means, cov, opac = e(b, v * r * samples, 3), e(b, v * r * samples, 3, 3), e(b, v * r * samples)
        csh, fsh = e(b, v * r * samples, d_color), e(b, v * r * samples, d_feature)
        index = e(G, dtype=torch.int32)
        args = _capi.LsGaussianHead(rays, r, width, height, samples, 32, d_color, d_feature, int(u is None), scale_min, scale_max,
                                    opacity_exponent, inv_gpp, dlog.data_ptr(), raw.data_ptr(), None if u is None else u.data_ptr(),
                                    ext.data_ptr(), intr.data_ptr(), near.data_ptr(), far.data_ptr())
        out = _capi.LsGaussianHeadOut(means.data_ptr(), cov.data_ptr(), opac.data_ptr(), csh.data_ptr(), fsh.data_ptr(),
                                      index.data_ptr())
        with torch.cuda.device(dev):
            _capi.check(_capi.load().ls_gaussian_head_forward(C.byref(args), C.byref(out), torch.cuda.current_stream().cuda_stream),
                        "ls_gaussian_head_forward")
        _capi.KERNEL_LAUNCHES[0] += 1
        ctx.save_for_backward(dlog, raw, u, ext, intr, near, far, index)
        ctx.cfg = (rays, r, width, height, samples, d_color, d_feature, scale_min, scale_max, opacity_exponent, inv_gpp)
        ctx.mark_non_differentiable(index)
        return means, cov, opac, csh, fsh, index

    @staticmethod
    def backward(ctx, g_means, g_cov, g_opac, g_csh, g_fsh, _g_index):
        dlog, raw, u, ext, intr, near, far, index = ctx.saved_tensors
        rays, r, width, height, samples, d_color, d_feature, scale_min, scale_max, opacity_exponent, inv_gpp = ctx.cfg
        z = lambda g, like_shape: (torch.zeros(like_shape, device=dlog.device) if g is None else g.contiguous())
        G = rays * samples
        g_means, g_cov, g_opac = z(g_means, (G, 3)), z(g_cov, (G, 3, 3)), z(g_opac, (G,))
        g_csh, g_fsh = z(g_csh, (G, d_color)), z(g_fsh, (G, d_feature))
        d_dlog, d_raw = torch.empty_like(dlog), torch.empty_like(raw)
        args = _capi.LsGaussianHead(rays, r, width, height, samples, 32, d_color, d_feature, int(u is None), scale_min, scale_max,
                                    opacity_exponent, inv_gpp, dlog.data_ptr(), raw.data_ptr(), None if u is None else u.data_ptr(),
                                    ext.data_ptr(), intr.data_ptr(), near.data_ptr(), far.data_ptr())
        grads = _capi.LsGaussianHeadGrad(index.data_ptr(), g_means.data_ptr(), g_cov.data_ptr(), g_opac.data_ptr(), g_csh.data_ptr(),
                                         g_fsh.data_ptr(), d_dlog.data_ptr(), d_raw.data_ptr())
        with torch.cuda.device(dlog.device):
            _capi.check(_capi.load().ls_gaussian_head_backward(C.byref(args), C.byref(grads), torch.cuda.current_stream().cuda_stream),
                        "ls_gaussian_head_backward")
        _capi.KERNEL_LAUNCHES[0] += 1
        return (d_dlog, d_raw) + (None,) * 14


def gaussian_head(dlog: Tensor, raw: Tensor, u: Optional[Tensor], extrinsics: Tensor, intrinsics: Tensor, near: Tensor,
                  far: Tensor, image_shape: tuple[int, int], samples: int, d_color: int, d_feature: int, scale_min: float,
                  scale_max: float, opacity_exponent: float, gaussians_per_pixel: int):
    """dlog (b, v, r, 64), raw (b, v, r, 9 + d_color + d_feature), u (b, v, r, samples) uniform draws or None (top-1),
    cameras (b, v, ...)  ->  means (b, G, 3), covariances (b, G, 3, 3), opacities (b, G), colour SH rows (b, G, d_color),
    feature SH rows (b, G, d_feature), sampled bucket index (b*G) with G = v * r * samples ordered (view, ray, sample)."""
    h, w = image_shape
    return _GaussianHead.apply(dlog, raw, u, extrinsics, intrinsics, near, far, h, w, samples, d_color, d_feature,
                               float(scale_min), float(scale_max), float(opacity_exponent), 1.0 / gaussians_per_pixel)
