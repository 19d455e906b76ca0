"""`Conv2d` / `ConvTranspose2d` on our sm_100a implicit-GEMM kernels (libls_raster.so, include/ls_conv.h).

Drop-in `nn.Conv2d` / `nn.ConvTranspose2d` (same parameters `weight`, `bias`, same shapes, so reference checkpoints
load).  On CUDA fp32 tensors the forward, input-gradient and weight-gradient passes run as TMA -> tcgen05 (TF32)
implicit GEMMs over NHWC memory: tensors keep their logical NCHW shape and live in torch's `channels_last` memory
format, which IS (N, H, W, C); weights in channels_last are (Cout, R, S, Cin), the K-major matrix the kernel wants
(`module.to(memory_format=torch.channels_last)` once avoids a per-call copy).  Bias and the following ReLU / GELU /
SiLU / LeakyReLU(0.2) are fused into the epilogue (`act=`); channel counts that are not multiples of 4 (RGB images,
the 7-channel skip, 1-channel logits) are zero-padded around the call.  The reference gets all of these from cuDNN
(/root/reference/src/model/encoder/epipolar/epipolar_transformer.py:68-74, autoencoder_kl.py:93-124,
discriminator_patch_gan.py:28-103).  CPU tensors (host-logic tests) and unsupported configurations (groups,
dilation, overlapping transposed convolutions) take torch's own path.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import torch
import torch.nn.functional as F
from torch import Tensor, nn

from . import _capi

ENABLED = True      # set False for A/B comparisons against cuDNN
ACT = {"none": _capi.ACT_NONE, "relu": _capi.ACT_RELU, "gelu": _capi.ACT_GELU, "silu": _capi.ACT_SILU,
       "lrelu": _capi.ACT_LRELU}
CL = torch.channels_last


def _cl(t: Tensor) -> Tensor:
    """Dense NHWC memory (a no-op for tensors that already are channels_last)."""
    return t.contiguous(memory_format=CL)


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def act_backward(gy: Tensor, saved: Tensor, act: str) -> Tensor:
    """gy * act'(saved) elementwise on our kernel; `saved` is the pre-activation (for relu / lrelu the output works too)."""
    out = torch.empty_like(gy)
    with torch.cuda.device(gy.device):
        _capi.check(_capi.load().ls_act_backward(gy.data_ptr(), saved.data_ptr(), out.data_ptr(), gy.numel(), ACT[act],
                                                 _stream()), "ls_act_backward")
    _capi.KERNEL_LAUNCHES[0] += 1
    return out


class _ConvFn(torch.autograd.Function):
    """x (N, Cin, H, W), weight (Cout, Cin, R, S) [transposed: (Cin, Cout, R, S)], both read as channels_last."""

    @staticmethod
    @torch.amp.custom_fwd(device_type="cuda", cast_inputs=torch.float32)
    def forward(ctx, x: Tensor, weight: Tensor, bias: Optional[Tensor], stride: int, pad: int, act: str, transposed: bool):
        x, w = _cl(x), _cl(weight)
        N, Cin, H, W = x.shape
        Cout = w.shape[1] if transposed else w.shape[0]
        R, S = w.shape[2:]
        desc = _capi.LsConv2d(N, H, W, Cin, Cout, R, S, stride, pad, int(transposed))
        oh, ow = C.c_int32(), C.c_int32()
        lib = _capi.load()
        _capi.check(lib.ls_conv2d_out_size(C.byref(desc), C.byref(oh), C.byref(ow)), "ls_conv2d_out_size")
        y = torch.empty((N, Cout, oh.value, ow.value), dtype=torch.float32, device=x.device, memory_format=CL)
        keep_pre = act in ("gelu", "silu")
        pre = torch.empty_like(y) if keep_pre else None
        with torch.cuda.device(x.device):
            _capi.check(lib.ls_conv2d_forward(C.byref(desc), x.data_ptr(), w.data_ptr(),
                                              None if bias is None else bias.data_ptr(), y.data_ptr(),
                                              None if pre is None else pre.data_ptr(), ACT[act], _stream()), "ls_conv2d_forward")
        _capi.KERNEL_LAUNCHES[0] += R * S if transposed else 1
        # algorithmic MACs: every (output pixel, Cout, tap, Cin) of a convolution = every (input pixel, ...) of a transposed one
        ctx.flops = 2.0 * N * (H * W if transposed else oh.value * ow.value) * Cout * R * S * Cin
        _capi.FLOPS["conv"] += ctx.flops
        ctx.desc, ctx.act, ctx.has_bias = desc, act, bias is not None
        ctx.save_for_backward(x, w, pre if keep_pre else (y if act != "none" else None))
        return y

    @staticmethod
    @torch.amp.custom_bwd(device_type="cuda")
    def backward(ctx, gy: Tensor):
        x, w, saved = ctx.saved_tensors
        desc, lib = ctx.desc, _capi.load()
        gy = _cl(gy)
        if ctx.act != "none":
            gy = act_backward(gy, saved, ctx.act)
        gx = gw = gb = None
        with torch.cuda.device(x.device):
            if ctx.needs_input_grad[0]:
                gx = torch.empty_like(x)                      # preserves channels_last
                _capi.check(lib.ls_conv2d_dgrad(C.byref(desc), gy.data_ptr(), w.data_ptr(), gx.data_ptr(), _stream()),
                            "ls_conv2d_dgrad")
                _capi.KERNEL_LAUNCHES[0] += 1 if desc.transposed else desc.stride * desc.stride
                _capi.FLOPS["conv"] += ctx.flops
            if ctx.needs_input_grad[1]:
                gw = torch.empty_like(w)
                _capi.check(lib.ls_conv2d_wgrad(C.byref(desc), gy.data_ptr(), x.data_ptr(), gw.data_ptr(), _stream()),
                            "ls_conv2d_wgrad")
                _capi.KERNEL_LAUNCHES[0] += 1
                _capi.FLOPS["conv"] += ctx.flops
        if ctx.has_bias and ctx.needs_input_grad[2]:
            from .gemm import col_sum
            gb = col_sum(gy.permute(0, 2, 3, 1).reshape(-1, gy.shape[1]))
        return gx, gw, gb, None, None, None, None


class _UpConvFn(torch.autograd.Function):
    """conv3x3(upsample_nearest_2x(x)) + bias without the up-sampled tensor: four 2x2 convolutions with folded weights on the
    low-resolution input (ls_upconv2x_*, include/ls_conv.h).  x (N, C, H, W) -> (N, Cout, 2H, 2W)."""

    @staticmethod
    @torch.amp.custom_fwd(device_type="cuda", cast_inputs=torch.float32)
    def forward(ctx, x: Tensor, weight: Tensor, bias: Optional[Tensor]):
        x, w = _cl(x), _cl(weight)
        N, Cin, H, W = x.shape
        Cout = w.shape[0]
        desc = _capi.LsConv2d(N, H, W, Cin, Cout, 3, 3, 1, 1, 0)
        lib = _capi.load()
        n = C.c_int64()
        _capi.check(lib.ls_upconv2x_workspace(C.byref(desc), C.byref(n)), "ls_upconv2x_workspace")
        wk = torch.empty(n.value, dtype=torch.float32, device=x.device)
        y = torch.empty((N, Cout, 2 * H, 2 * W), dtype=torch.float32, device=x.device, memory_format=CL)
        with torch.cuda.device(x.device):
            _capi.check(lib.ls_upconv2x_forward(C.byref(desc), x.data_ptr(), w.data_ptr(), None if bias is None else bias.data_ptr(),
                                                y.data_ptr(), wk.data_ptr(), _stream()), "ls_upconv2x_forward")
        _capi.KERNEL_LAUNCHES[0] += 5
        ctx.flops = 2.0 * N * H * W * 16 * Cout * Cin       # executed: 4 parity classes x 2x2 folded taps (the unfolded form does 36)
        _capi.FLOPS["conv"] += ctx.flops
        ctx.desc, ctx.has_bias = desc, bias is not None
        ctx.save_for_backward(x, w, wk)
        return y

    @staticmethod
    @torch.amp.custom_bwd(device_type="cuda")
    def backward(ctx, gy: Tensor):
        x, w, wk = ctx.saved_tensors
        desc, lib = ctx.desc, _capi.load()
        gy = _cl(gy)
        gx = gw = gb = None
        with torch.cuda.device(x.device):
            if ctx.needs_input_grad[0]:
                gx = torch.empty_like(x)
                _capi.check(lib.ls_upconv2x_dgrad(C.byref(desc), gy.data_ptr(), wk.data_ptr(), gx.data_ptr(), _stream()),
                            "ls_upconv2x_dgrad")
                _capi.KERNEL_LAUNCHES[0] += 1
                _capi.FLOPS["conv"] += ctx.flops
            if ctx.needs_input_grad[1]:
                gw = torch.empty_like(w)
                scratch = torch.empty(16 * desc.Cout * desc.Cin, dtype=torch.float32, device=x.device)
                _capi.check(lib.ls_upconv2x_wgrad(C.byref(desc), gy.data_ptr(), x.data_ptr(), gw.data_ptr(), scratch.data_ptr(),
                                                  _stream()), "ls_upconv2x_wgrad")
                _capi.KERNEL_LAUNCHES[0] += 5
                _capi.FLOPS["conv"] += ctx.flops
        if ctx.has_bias and ctx.needs_input_grad[2]:
            from .gemm import col_sum
            gb = col_sum(gy.permute(0, 2, 3, 1).reshape(-1, gy.shape[1]))
        return gx, gw, gb


def upsample2x_conv3x3(x: Tensor, weight: Tensor, bias: Optional[Tensor]) -> Tensor:
    """F.conv2d(F.interpolate(x, scale_factor=2, mode="nearest"), weight, bias, padding=1) on the folded-weight kernels."""
    if not x.is_cuda:
        raise RuntimeError("upsample2x_conv3x3 needs CUDA tensors: latentsplat_b200 has no CPU fallback")
    return _UpConvFn.apply(x, weight, bias)


def _pad_channels(t: Tensor, dim: int, n: int) -> Tensor:
    if n == 0:
        return t
    pad = [0, 0] * (t.dim() - 1 - dim) + [0, n]
    return F.pad(t, pad)


def conv2d(x: Tensor, weight: Tensor, bias: Optional[Tensor], stride: int, pad: int, act: str = "none",
           transposed: bool = False) -> Tensor:
    """F.conv2d / F.conv_transpose2d (+ bias + activation) on the implicit-GEMM kernels.  CUDA fp32 only."""
    if not x.is_cuda:
        raise RuntimeError("conv2d needs CUDA tensors: latentsplat_b200 has no CPU fallback")
    cin_dim, cout_dim = (0, 1) if transposed else (1, 0)
    Cin, Cout = weight.shape[cin_dim], weight.shape[cout_dim]
    pc, po = (-Cin) % 4, (-Cout) % 4
    if pc:
        x, weight = _pad_channels(x, 1, pc), _pad_channels(weight, cin_dim, pc)
    if po:
        weight = _pad_channels(weight, cout_dim, po)
        bias = None if bias is None else F.pad(bias, (0, po))
    y = _ConvFn.apply(x, weight, bias, stride, pad, act, transposed)
    return y[:, :Cout] if po else y


def _int_pair(v) -> Optional[int]:
    if isinstance(v, int):
        return v
    if isinstance(v, (tuple, list)) and len(v) == 2 and v[0] == v[1]:
        return int(v[0])
    return None


def _apply_act(y: Tensor, act: str) -> Tensor:
    if act == "relu":
        return F.relu(y)
    if act == "gelu":
        return F.gelu(y)
    if act == "silu":
        return F.silu(y)
    if act == "lrelu":
        return F.leaky_relu(y, 0.2)
    return y


class Conv2d(nn.Conv2d):
    """nn.Conv2d with an optional fused activation (`act`: "none" | "relu" | "gelu" | "silu" | "lrelu")."""

    def __init__(self, *args, act: str = "none", **kwargs):
        super().__init__(*args, **kwargs)
        assert act in ACT
        self.act = act

    def _native(self, x: Tensor) -> bool:
        return (ENABLED and x.is_cuda and x.dtype == torch.float32 and x.dim() == 4 and x.shape[0] > 0 and self.groups == 1
                and self.padding_mode == "zeros" and not isinstance(self.padding, str) and tuple(self.dilation) == (1, 1)
                and _int_pair(self.stride) is not None and _int_pair(self.padding) is not None
                and 1 <= _int_pair(self.stride) <= 8 and self.weight.dtype == torch.float32)

    def forward(self, input: Tensor) -> Tensor:
        if self._native(input):
            return conv2d(input, self.weight, self.bias, _int_pair(self.stride), _int_pair(self.padding), self.act)
        return _apply_act(super().forward(input), self.act)

    def forward_upsampled2x(self, input: Tensor) -> Tensor:
        """self(F.interpolate(input, scale_factor=2, mode="nearest")) -- fused (no up-sampled tensor, 2.25x fewer flops) for the
        3x3 / stride 1 / pad 1 case on CUDA, the explicit two-step sequence otherwise."""
        if (self._native(input) and self.act == "none" and tuple(self.kernel_size) == (3, 3) and _int_pair(self.stride) == 1
                and _int_pair(self.padding) == 1 and self.in_channels % 4 == 0 and self.out_channels % 4 == 0):
            return upsample2x_conv3x3(input, self.weight, self.bias)
        return self.forward(F.interpolate(input, scale_factor=2.0, mode="nearest"))


class ConvTranspose2d(nn.ConvTranspose2d):
    """nn.ConvTranspose2d; the non-overlapping case (kernel == stride, no padding: the 4x up-scalers of
    epipolar_transformer.py:69 and image_self_attention.py:52) runs on the implicit-GEMM kernels."""

    def __init__(self, *args, act: str = "none", **kwargs):
        super().__init__(*args, **kwargs)
        assert act in ACT
        self.act = act

    def _native(self, x: Tensor) -> bool:
        k, s = _int_pair(self.kernel_size), _int_pair(self.stride)
        return (ENABLED and x.is_cuda and x.dtype == torch.float32 and x.dim() == 4 and x.shape[0] > 0 and self.groups == 1
                and k is not None and k == s and 1 <= s <= 8 and _int_pair(self.padding) == 0
                and _int_pair(self.output_padding) == 0 and tuple(self.dilation) == (1, 1) and self.weight.dtype == torch.float32)

    def forward(self, input: Tensor, output_size=None) -> Tensor:
        if output_size is None and self._native(input):
            return conv2d(input, self.weight, self.bias, _int_pair(self.stride), 0, self.act, transposed=True)
        return _apply_act(super().forward(input, output_size), self.act)
