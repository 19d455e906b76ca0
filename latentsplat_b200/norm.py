"""GroupNorm (+ SiLU) on our sm_100a kernels (libls_raster.so, include/ls_norm.h).

`GroupNorm` is a drop-in `nn.GroupNorm` (same parameters, so diffusers / latentSplat checkpoints load) with an `act`
attribute ("none" | "silu"): the VAE decoder's `nonlinearity(norm(x))` pairs become one call.  CUDA fp32 NCHW tensors
take the fused kernels; anything else (CPU host-logic tests) takes `F.group_norm` (+ `F.silu`)."""
from __future__ import annotations

import ctypes as C

import torch
import torch.nn.functional as F
from torch import Tensor, nn

from . import _capi

ENABLED = True      # set False for A/B comparisons against torch's GroupNorm + SiLU


class _GroupNormFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x: Tensor, weight: Tensor, bias: Tensor, groups: int, eps: float, act: int):
        x = x.contiguous()
        N, Cn = x.shape[:2]
        hw = x[0, 0].numel()
        stats = torch.empty((N, groups, 2), dtype=torch.float64, device=x.device)
        y = torch.empty_like(x)
        a = _capi.LsGroupNorm(N, Cn, groups, act, hw, eps, x.data_ptr(), weight.data_ptr(), bias.data_ptr(), stats.data_ptr())
        with torch.cuda.device(x.device):
            _capi.check(_capi.load().ls_groupnorm_forward(C.byref(a), y.data_ptr(), torch.cuda.current_stream().cuda_stream),
                        "ls_groupnorm_forward")
        _capi.KERNEL_LAUNCHES[0] += 2
        ctx.save_for_backward(x, weight, bias, stats)
        ctx.cfg = (groups, eps, act)
        return y

    @staticmethod
    def backward(ctx, dy: Tensor):
        x, weight, bias, stats = ctx.saved_tensors
        groups, eps, act = ctx.cfg
        dy = dy.contiguous()
        N, Cn = x.shape[:2]
        sums = torch.empty((N, Cn, 2), dtype=torch.float64, device=x.device)
        dx = torch.empty_like(x)
        a = _capi.LsGroupNorm(N, Cn, groups, act, x[0, 0].numel(), eps, x.data_ptr(), weight.data_ptr(), bias.data_ptr(),
                              stats.data_ptr())
        with torch.cuda.device(x.device):
            _capi.check(_capi.load().ls_groupnorm_backward(C.byref(a), dy.data_ptr(), dx.data_ptr(), sums.data_ptr(),
                                                           torch.cuda.current_stream().cuda_stream), "ls_groupnorm_backward")
        _capi.KERNEL_LAUNCHES[0] += 2
        per_channel = sums.sum(dim=0).to(torch.float32)            # (C, 2): d beta, d gamma
        return dx, per_channel[:, 1].contiguous(), per_channel[:, 0].contiguous(), None, None, None


class _GroupNormNhwcFn(torch.autograd.Function):
    """GroupNorm (+ SiLU) over a token-major tensor x (N, L, C) -- the memory of an NCHW tensor in channels_last format."""

    @staticmethod
    def forward(ctx, x: Tensor, weight: Tensor, bias: Tensor, groups: int, eps: float, act: int):
        N, L, Cn = x.shape
        stats = torch.empty((N, Cn, 2), dtype=torch.float64, device=x.device)
        y = torch.empty_like(x)
        a = _capi.LsGroupNorm(N, Cn, groups, act, L, eps, x.data_ptr(), weight.data_ptr(), bias.data_ptr(), stats.data_ptr())
        with torch.cuda.device(x.device):
            _capi.check(_capi.load().ls_groupnorm_nhwc_forward(C.byref(a), y.data_ptr(), torch.cuda.current_stream().cuda_stream),
                        "ls_groupnorm_nhwc_forward")
        _capi.KERNEL_LAUNCHES[0] += 2
        ctx.save_for_backward(x, weight, bias, stats)
        ctx.cfg = (groups, eps, act)
        return y

    @staticmethod
    def backward(ctx, dy: Tensor):
        x, weight, bias, stats = ctx.saved_tensors
        groups, eps, act = ctx.cfg
        N, L, Cn = x.shape
        dy = dy.contiguous()
        sums = torch.empty((N, Cn, 2), dtype=torch.float64, device=x.device)
        dx = torch.empty_like(x)
        a = _capi.LsGroupNorm(N, Cn, groups, act, L, eps, x.data_ptr(), weight.data_ptr(), bias.data_ptr(), stats.data_ptr())
        with torch.cuda.device(x.device):
            _capi.check(_capi.load().ls_groupnorm_nhwc_backward(C.byref(a), dy.data_ptr(), dx.data_ptr(), sums.data_ptr(),
                                                                torch.cuda.current_stream().cuda_stream), "ls_groupnorm_nhwc_backward")
        _capi.KERNEL_LAUNCHES[0] += 2
        per_channel = sums.sum(dim=0).to(torch.float32)            # (C, 2): d beta, d gamma
        return dx, per_channel[:, 1].contiguous(), per_channel[:, 0].contiguous(), None, None, None


def group_norm_tokens(t: Tensor, groups: int, weight: Tensor, bias: Tensor, eps: float, act: str = "none") -> Tensor:
    """GroupNorm of a token-major tensor t (N, L, C) == F.group_norm(t.transpose(1, 2), ...).transpose(1, 2)."""
    Cn = t.shape[-1]
    if (ENABLED and t.is_cuda and t.dtype == torch.float32 and t.dim() == 3 and Cn % 4 == 0 and Cn <= 1024
            and weight is not None and bias is not None and t.shape[0] <= 65535 and t.numel() > 0):
        return _GroupNormNhwcFn.apply(t.contiguous(), weight, bias, groups, eps, 1 if act == "silu" else 0)
    y = F.group_norm(t.transpose(1, 2), groups, weight, bias, eps).transpose(1, 2)
    return F.silu(y) if act == "silu" else y


def group_norm(x: Tensor, groups: int, weight: Tensor, bias: Tensor, eps: float, act: str = "none") -> Tensor:
    if (ENABLED and x.is_cuda and x.dtype == torch.float32 and x.dim() == 4 and not x.is_contiguous()
            and x.is_contiguous(memory_format=torch.channels_last) and x.shape[1] % 4 == 0 and x.shape[1] <= 1024
            and weight is not None and bias is not None and x.shape[0] <= 65535):
        # channels_last activations (the implicit-GEMM convolutions' layout): normalise the (N, H*W, C) memory in place
        N, Cn, H, W = x.shape
        y = _GroupNormNhwcFn.apply(x.permute(0, 2, 3, 1).reshape(N, H * W, Cn), weight, bias, groups, eps, 1 if act == "silu" else 0)
        return y.view(N, H, W, Cn).permute(0, 3, 1, 2)
    if (ENABLED and x.is_cuda and x.dtype == torch.float32 and x.dim() >= 3 and x[0, 0].numel() % 4 == 0
            and weight is not None and bias is not None and x.shape[0] * x.shape[1] <= 65535):
        return _GroupNormFn.apply(x, weight, bias, groups, eps, 1 if act == "silu" else 0)
    y = F.group_norm(x, groups, weight, bias, eps)
    return F.silu(y) if act == "silu" else y


class GroupNorm(nn.GroupNorm):
    """nn.GroupNorm with an optional fused activation (`act="silu"`); parameters `weight`, `bias` as nn.GroupNorm."""

    def __init__(self, num_groups: int, num_channels: int, eps: float = 1e-5, affine: bool = True, act: str = "none"):
        super().__init__(num_groups, num_channels, eps=eps, affine=affine)
        assert act in ("none", "silu")
        self.act = act

    def forward(self, input: Tensor) -> Tensor:
        return group_norm(input, self.num_groups, self.weight, self.bias, self.eps, self.act)


# ------------------------------------------------------------------------------------------------------------------
# LayerNorm
# ------------------------------------------------------------------------------------------------------------------
class _LayerNormFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x: Tensor, weight: Tensor, bias: Tensor, eps: float):
        shape = x.shape
        x2 = x.reshape(-1, shape[-1]).contiguous()
        rows, Cn = x2.shape
        y = torch.empty_like(x2)
        mean_rstd = torch.empty((rows, 2), dtype=torch.float32, device=x.device)
        with torch.cuda.device(x.device):
            _capi.check(_capi.load().ls_layernorm_forward(x2.data_ptr(), weight.data_ptr(), bias.data_ptr(), y.data_ptr(),
                                                          mean_rstd.data_ptr(), rows, Cn, eps,
                                                          torch.cuda.current_stream().cuda_stream), "ls_layernorm_forward")
        _capi.KERNEL_LAUNCHES[0] += 1
        ctx.save_for_backward(x2, weight, mean_rstd)
        return y.view(shape)

    @staticmethod
    def backward(ctx, dy: Tensor):
        x2, weight, mean_rstd = ctx.saved_tensors
        rows, Cn = x2.shape
        dy2 = dy.reshape(rows, Cn).contiguous()
        dx = torch.empty_like(x2)
        dgb = torch.zeros((2, Cn), dtype=torch.float32, device=x2.device)
        with torch.cuda.device(x2.device):
            _capi.check(_capi.load().ls_layernorm_backward(x2.data_ptr(), dy2.data_ptr(), weight.data_ptr(), mean_rstd.data_ptr(),
                                                           dx.data_ptr(), dgb[0].data_ptr(), dgb[1].data_ptr(), rows, Cn,
                                                           torch.cuda.current_stream().cuda_stream), "ls_layernorm_backward")
        _capi.KERNEL_LAUNCHES[0] += 1
        return dx.view(dy.shape), dgb[0], dgb[1], None


def layer_norm(x: Tensor, weight: Tensor, bias: Tensor, eps: float) -> Tensor:
    Cn = x.shape[-1]
    if (ENABLED and x.is_cuda and x.dtype == torch.float32 and weight is not None and bias is not None
            and Cn % 128 == 0 and Cn <= 1024 and x.numel() > 0):
        return _LayerNormFn.apply(x, weight, bias, eps)
    return F.layer_norm(x, (Cn,), weight, bias, eps)


class LayerNorm(nn.LayerNorm):
    """nn.LayerNorm over the last dimension on our kernel when CUDA fp32 and C % 128 == 0 (parameters as nn.LayerNorm)."""

    def forward(self, input: Tensor) -> Tensor:
        if len(self.normalized_shape) != 1:
            return super().forward(input)
        return layer_norm(input, self.weight, self.bias, self.eps)
