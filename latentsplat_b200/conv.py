"""`Conv2d`: nn.Conv2d whose bias add and bias gradient run on our kernels (libls_raster.so, include/ls_norm.h).

The convolution itself stays a library call (cuDNN TF32, as in the reference); what changes is its epilogue: torch adds
the bias with a broadcasting elementwise kernel and reduces its gradient with a strided reduction (6 + 5 ms of the
step at the bench shape), here both are plane-wise float4 passes at HBM speed.  Same parameters as nn.Conv2d
(`weight`, `bias`), so reference checkpoints load; CPU tensors and unusual configurations take nn.Conv2d's own path."""
from __future__ import annotations

import torch
from torch import Tensor, nn

from . import _capi

ENABLED = True      # set False for A/B comparisons against torch's own bias handling


class _ConvBiasFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x: Tensor, weight: Tensor, bias: Tensor, stride, padding, dilation, groups: int):
        y = torch.ops.aten.convolution(x, weight, None, stride, padding, dilation, False, [0, 0], groups)
        if y.is_contiguous():
            N, Cn = y.shape[:2]
            with torch.cuda.device(x.device):
                _capi.check(_capi.load().ls_conv_bias_add(y.data_ptr(), bias.data_ptr(), N, Cn, y[0, 0].numel(),
                                                          torch.cuda.current_stream().cuda_stream), "ls_conv_bias_add")
            _capi.KERNEL_LAUNCHES[0] += 1
        else:       # channels-last output (the input arrived channels-last): the bias is the fastest-varying axis, torch's
            y.add_(bias.view(1, -1, 1, 1))      # vectorised row-broadcast add is already right; never force a layout copy
        ctx.save_for_backward(x, weight)
        ctx.cfg = (stride, padding, dilation, groups)
        return y

    @staticmethod
    def backward(ctx, gy: Tensor):
        x, weight = ctx.saved_tensors
        stride, padding, dilation, groups = ctx.cfg
        gx, gw, _ = torch.ops.aten.convolution_backward(gy, x, weight, None, stride, padding, dilation, False, [0, 0], groups,
                                                        [ctx.needs_input_grad[0], ctx.needs_input_grad[1], False])
        gb = None
        if ctx.needs_input_grad[2] and not gy.is_contiguous():
            if gy.is_contiguous(memory_format=torch.channels_last):      # memory is a row-major (N*H*W, C) matrix
                from .gemm import col_sum
                gb = col_sum(gy.permute(0, 2, 3, 1).reshape(-1, gy.shape[1]))
            else:
                gb = gy.sum(dim=(0, 2, 3))
        elif ctx.needs_input_grad[2]:
            N, Cn = gy.shape[:2]
            gb = torch.zeros(Cn, dtype=torch.float32, device=gy.device)
            with torch.cuda.device(gy.device):
                _capi.check(_capi.load().ls_conv_bias_grad(gy.data_ptr(), gb.data_ptr(), N, Cn, gy[0, 0].numel(),
                                                           torch.cuda.current_stream().cuda_stream), "ls_conv_bias_grad")
            _capi.KERNEL_LAUNCHES[0] += 1
        return gx, gw, gb, None, None, None, None


class Conv2d(nn.Conv2d):
    def forward(self, input: Tensor) -> Tensor:
        if (ENABLED and input.is_cuda and input.dtype == torch.float32 and self.bias is not None and input.dim() == 4
                and self.padding_mode == "zeros" and not isinstance(self.padding, str)
                and input.shape[0] * self.out_channels <= 65535 and input.shape[0] > 0):
            return _ConvBiasFn.apply(input, self.weight, self.bias, list(self.stride), list(self.padding), list(self.dilation),
                                     self.groups)
        return super().forward(input)
