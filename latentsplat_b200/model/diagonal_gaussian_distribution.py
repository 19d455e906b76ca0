"""Diagonal Gaussian over arbitrary-shaped tensors.

Same interface and numerics as /root/reference/src/model/diagonal_gaussian_distribution.py:8-95
(clamped log-variance, `sample/kl/nll/mode`, `params` = cat(mean, logvar) along `dim`).
"""
from __future__ import annotations

import math
from typing import Optional, Tuple

import torch
from torch import Tensor


def _fused_ok(params: Optional[Tensor], dim: int) -> bool:
    if params is None or not (params.is_cuda and params.dtype == torch.float32 and params.is_contiguous()):
        return False
    width = 1
    for n in params.shape[dim:]:
        width *= n
    return width % 8 == 0 and params.numel() > 0


class _ReparamFn(torch.autograd.Function):
    """sample = mean + std * eps on libls_raster.so::ls_reparam_* (include/ls_ghead.h): params = cat(mean, logvar) along `dim`, i.e.
    rows of [mean | logvar] once everything from `dim` on is flattened."""

    @staticmethod
    def forward(ctx, params: Tensor, eps: Tensor, dim: int, lo: float, hi: float) -> Tensor:
        import ctypes as C
        from latentsplat_b200 import _capi
        dim = dim % params.dim()
        width = params[(0,) * dim].numel() if dim else params.numel()
        rows, half = params.numel() // width, width // 2
        eps = eps.contiguous()
        out = torch.empty_like(eps)
        with torch.cuda.device(params.device):
            _capi.check(_capi.load().ls_reparam_forward(params.data_ptr(), eps.data_ptr(), out.data_ptr(), rows, half, lo, hi,
                                                        torch.cuda.current_stream().cuda_stream), "ls_reparam_forward")
        _capi.KERNEL_LAUNCHES[0] += 1
        ctx.save_for_backward(params, eps)
        ctx.cfg = (rows, half, lo, hi)
        return out

    @staticmethod
    def backward(ctx, g: Tensor):
        from latentsplat_b200 import _capi
        params, eps = ctx.saved_tensors
        rows, half, lo, hi = ctx.cfg
        g = g.contiguous()
        d_params = torch.empty_like(params)
        with torch.cuda.device(params.device):
            _capi.check(_capi.load().ls_reparam_backward(params.data_ptr(), eps.data_ptr(), g.data_ptr(), d_params.data_ptr(), rows, half,
                                                         lo, hi, torch.cuda.current_stream().cuda_stream), "ls_reparam_backward")
        _capi.KERNEL_LAUNCHES[0] += 1
        return d_params, None, None, None, None


class DiagonalGaussianDistribution:
    def __init__(self, mean: Optional[Tensor] = None, logvar: Optional[Tensor] = None,
                 params: Optional[Tensor] = None, dim: int = 0,
                 logvar_interval: Tuple[float, float] = (-30.0, 20.0)):
        if params is not None and (mean is not None or logvar is not None):
            raise AssertionError("If params are given, mean and logvar are not expected")
        if mean is None and params is None:
            raise AssertionError("Either mean or params must be given")
        self.logvar_interval = logvar_interval
        self.dim = dim
        self._params = None
        self.mean = mean
        self._set_logvar(logvar)
        if params is not None:
            self.params = params

    # -- logvar / std / var ------------------------------------------------------------
    def _set_logvar(self, val: Optional[Tensor]) -> None:
        self._raw_logvar = val
        if val is None:
            self._logvar, self._std, self._var = None, 0.0, 0.0  # zero variance by default
            return
        if val.shape != self.mean.shape:
            raise AssertionError("Shapes of mean and logvar must be identical")
        self._raw_logvar = val
        self._logvar = self._std = self._var = None            # clamp / exp run on first use (the render path needs none of them
                                                               # when it samples through the fused kernel below)

    @property
    def logvar(self) -> Optional[Tensor]:
        if self._logvar is None and self._raw_logvar is not None:
            self._logvar = torch.clamp(self._raw_logvar, *self.logvar_interval)
        return self._logvar

    @logvar.setter
    def logvar(self, val: Optional[Tensor]) -> None:
        self._set_logvar(val)

    @property
    def std(self):
        if self._std is None:
            self._std = torch.exp(0.5 * self.logvar)
        return self._std

    @property
    def var(self):
        if self._var is None:
            self._var = torch.exp(self.logvar)
        return self._var

    # -- params ------------------------------------------------------------------------
    @property
    def params(self) -> Tensor:
        if self._params is None:
            if self.logvar is None:
                raise AssertionError("Trying accessing params without params or logvar")
            return torch.cat((self.mean, self.logvar), dim=self.dim)
        return self._params

    @params.setter
    def params(self, val: Optional[Tensor]) -> None:
        if val is not None:
            mean, logvar = val.chunk(2, dim=self.dim)
            self.mean = mean
            self._set_logvar(logvar)
        self._params = val

    @property
    def device(self) -> torch.device:
        return self.mean.device

    # -- distribution ops --------------------------------------------------------------
    def sample(self) -> Tensor:
        if self._raw_logvar is None:
            return self.mean
        eps = torch.randn_like(self.mean)
        if _fused_ok(self._params, self.dim):
            # CUDA: mean + exp(0.5 clamp(logvar)) * eps in one pass over the packed params, its gradient in one pass back
            return _ReparamFn.apply(self._params, eps, self.dim, *self.logvar_interval)
        return self.mean + self.std * eps

    def mode(self) -> Tensor:
        return self.mean

    def kl(self, other: Optional["DiagonalGaussianDistribution"] = None) -> Tensor:
        if self.logvar is None:
            return torch.zeros_like(self.mean)
        if other is None:
            return 0.5 * (self.mean ** 2 + self.var - 1.0 - self.logvar)
        return 0.5 * ((self.mean - other.mean) ** 2 / other.var + self.var / other.var - 1.0
                      - self.logvar + other.logvar)

    def nll(self, sample: Tensor) -> Tensor:
        if self.logvar is None:
            return torch.zeros_like(self.mean, device=self.device)
        return 0.5 * (math.log(2.0 * math.pi) + self.logvar + (sample - self.mean) ** 2 / self.var)
