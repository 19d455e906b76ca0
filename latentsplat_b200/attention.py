"""Fused single-query attention (libls_raster.so::ls_sq_attention_*, include/ls_gemm.h) as an autograd function."""
from __future__ import annotations

import torch
from torch import Tensor

from . import _capi


class _SingleQueryAttention(torch.autograd.Function):
    @staticmethod
    def forward(ctx, q: Tensor, kv: Tensor, heads: int, scale: float):
        R, HD = q.shape
        S = kv.shape[1]
        q, kv = q.contiguous(), kv.contiguous()
        out = torch.empty_like(q)
        p = torch.empty((R, heads, S), dtype=torch.float32, device=q.device)
        with torch.cuda.device(q.device):
            _capi.check(_capi.load().ls_sq_attention_forward(q.data_ptr(), kv.data_ptr(), out.data_ptr(), p.data_ptr(), R,
                                                             heads, S, HD // heads, scale,
                                                             torch.cuda.current_stream().cuda_stream), "ls_sq_attention_forward")
        _capi.KERNEL_LAUNCHES[0] += 1
        ctx.save_for_backward(q, kv, p)
        ctx.cfg = (heads, scale)
        return out

    @staticmethod
    def backward(ctx, dout: Tensor):
        q, kv, p = ctx.saved_tensors
        heads, scale = ctx.cfg
        R, HD = q.shape
        S = kv.shape[1]
        dout = dout.contiguous()
        dq, dkv = torch.empty_like(q), torch.empty_like(kv)
        with torch.cuda.device(q.device):
            _capi.check(_capi.load().ls_sq_attention_backward(q.data_ptr(), kv.data_ptr(), p.data_ptr(), dout.data_ptr(),
                                                              dq.data_ptr(), dkv.data_ptr(), R, heads, S, HD // heads, scale,
                                                              torch.cuda.current_stream().cuda_stream), "ls_sq_attention_backward")
        _capi.KERNEL_LAUNCHES[0] += 1
        return dq, dkv, None, None


def single_query_attention(q: Tensor, kv: Tensor, heads: int, scale: float) -> Tensor:
    """q (R, H*128), kv (R, S<=32, 2*H*128) = [K | V]  ->  softmax(q K^T * scale) V  as (R, H*128)."""
    return _SingleQueryAttention.apply(q, kv, heads, scale)


def supported(q: Tensor, kv: Tensor, heads: int) -> bool:
    return (q.is_cuda and q.dtype == torch.float32 and kv.dtype == torch.float32 and q.dim() == 2 and kv.dim() == 3
            and q.shape[1] == heads * 128 and kv.shape[2] == 2 * heads * 128 and kv.shape[1] <= 32)
