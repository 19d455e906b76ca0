"""Diagonal Gaussian over arbitrary-shaped tensors.

Same interface and numerics as /root/reference/src/model/diagonal_gaussian_distribution.py:8-95
(clamped log-variance, `sample/kl/nll/mode`, `params` = cat(mean, logvar) along `dim`).
"""
from __future__ import annotations

import math
from typing import Optional, Tuple

import torch
from torch import Tensor


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
        if val is None:
            self._logvar, self._std, self._var = None, 0.0, 0.0  # zero variance by default
            return
        if val.shape != self.mean.shape:
            raise AssertionError("Shapes of mean and logvar must be identical")
        self._logvar = torch.clamp(val, *self.logvar_interval)
        self._std = torch.exp(0.5 * self._logvar)
        self._var = torch.exp(self._logvar)

    @property
    def logvar(self) -> Optional[Tensor]:
        return self._logvar

    @logvar.setter
    def logvar(self, val: Optional[Tensor]) -> None:
        self._set_logvar(val)

    @property
    def std(self):
        return self._std

    @property
    def var(self):
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
        if isinstance(self.std, float) and self.std == 0:
            return self.mean
        return self.mean + self.std * torch.randn_like(self.mean)

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
