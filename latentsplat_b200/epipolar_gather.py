"""Fused epipolar-line feature gather + depth encoding (libls_raster.so, include/ls_epipolar.h).

One kernel instead of the reference's transpose -> grid_sample -> rearrange -> transpose -> mask multiply
(/root/reference/src/model/encoder/epipolar/epipolar_sampler.py:96-112) and the depth-encoding Linear + add
(epipolar_transformer.py:121-122).  CUDA only: there is no CPU fallback (the modules keep the explicit PyTorch
sequence for CPU host-logic tests and call this on CUDA tensors).
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import torch
from torch import Tensor

from . import _capi

ENABLED = True      # set False for A/B comparisons against the explicit grid_sample path


def _args(feat: Tensor, xy: Tensor, depth: Optional[Tensor], image: Tensor, valid: Tensor, width: int) -> _capi.LsEpipolarGather:
    rows, samples = xy.shape[0], xy.shape[1]
    return _capi.LsEpipolarGather(rows, samples, feat.shape[0], feat.shape[1], feat.shape[2], feat.shape[3], width,
                                  xy.data_ptr(), None if depth is None else depth.data_ptr(), image.data_ptr(), valid.data_ptr())


class _EpipolarGather(torch.autograd.Function):
    @staticmethod
    def forward(ctx, feat: Tensor, xy: Tensor, depth: Optional[Tensor], image: Tensor, valid: Tensor,
                weight: Optional[Tensor], bias: Optional[Tensor]):
        if not feat.is_cuda:
            raise RuntimeError("epipolar_gather needs CUDA tensors: latentsplat_b200 has no CPU fallback")
        feat, xy = feat.contiguous(), xy.contiguous()
        width = 0 if weight is None else weight.shape[1]
        if width:
            depth, weight, bias = depth.contiguous(), weight.contiguous(), bias.contiguous()
        assert feat.dtype == xy.dtype == valid.dtype == torch.float32 and image.dtype == torch.int32
        assert xy.dim() == 3 and xy.shape[2] == 2 and feat.dim() == 4 and image.shape == valid.shape == xy.shape[:1]
        z = torch.empty((*xy.shape[:2], feat.shape[3]), dtype=torch.float32, device=feat.device)
        a = _args(feat, xy, depth if width else None, image, valid, width)
        with torch.cuda.device(feat.device):
            _capi.check(_capi.load().ls_epipolar_gather_forward(C.byref(a), feat.data_ptr(), weight.data_ptr() if width else None,
                                                                bias.data_ptr() if width else None, z.data_ptr(),
                                                                torch.cuda.current_stream().cuda_stream), "ls_epipolar_gather_forward")
        _capi.KERNEL_LAUNCHES[0] += 1
        ctx.save_for_backward(xy, depth if width else None, image, valid)
        ctx.meta = (feat.shape, width)
        return z

    @staticmethod
    def backward(ctx, dz: Tensor):
        xy, depth, image, valid = ctx.saved_tensors
        shape, width = ctx.meta
        dz = dz.contiguous()
        dfeat = torch.zeros(shape, dtype=torch.float32, device=dz.device)
        dw = torch.zeros((shape[3], width), dtype=torch.float32, device=dz.device) if width else None
        db = torch.zeros((shape[3],), dtype=torch.float32, device=dz.device) if width else None
        a = _capi.LsEpipolarGather(xy.shape[0], xy.shape[1], shape[0], shape[1], shape[2], shape[3], width, xy.data_ptr(),
                                   depth.data_ptr() if width else None, image.data_ptr(), valid.data_ptr())
        with torch.cuda.device(dz.device):
            _capi.check(_capi.load().ls_epipolar_gather_backward(C.byref(a), dz.data_ptr(), dfeat.data_ptr(),
                                                                 dw.data_ptr() if width else None, db.data_ptr() if width else None,
                                                                 torch.cuda.current_stream().cuda_stream), "ls_epipolar_gather_backward")
        _capi.KERNEL_LAUNCHES[0] += 1
        return dfeat, None, None, None, None, dw, db


def epipolar_gather(feat: Tensor, xy: Tensor, depth: Optional[Tensor], image: Tensor, valid: Tensor,
                    weight: Optional[Tensor] = None, bias: Optional[Tensor] = None) -> Tensor:
    """feat (images, h, w, 128) channels-last features; xy (rows, samples<=32, 2) normalised coordinates; depth (rows, samples)
    relative disparity (only with weight/bias); image (rows) int32 feature-map index; valid (rows) 0/1 float;
    weight (128, 20), bias (128) = the depth-encoding Linear.  Returns z (rows, samples, 128)."""
    return _EpipolarGather.apply(feat, xy, depth, image, valid, weight, bias)


def supported(feat: Tensor, channels: int, samples: int, weight: Optional[Tensor]) -> bool:
    return (ENABLED and feat.is_cuda and feat.dtype == torch.float32 and channels == 128 and samples <= 32
            and (weight is None or tuple(weight.shape) == (128, 20)))
