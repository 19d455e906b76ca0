"""Multi-head self-attention core on our tcgen05 flash-attention kernels (libls_raster.so::ls_fmha_*, include/ls_fmha.h).

`attention_packed(qkv, heads, scale)`: qkv (B, L, 3*H*D) as produced by a fused qkv projection (DINO's `attn.qkv`, the
`to_qkv` of /root/reference/src/model/transformer/attention.py:45) -> softmax(q k^T scale) v as (B, L, H*D), reading q, k, v
in place (no head-major copies, no casts: fp32 in HBM, TF32 in the tensor core).  CUDA only.
"""
from __future__ import annotations

import ctypes as C

import torch
from torch import Tensor

from . import _capi

ENABLED = True      # set False for A/B comparisons against torch's SDPA


def supported(qkv: Tensor, heads: int) -> bool:
    if not (ENABLED and qkv.is_cuda and qkv.dtype == torch.float32 and qkv.dim() == 3 and qkv.shape[-1] % (3 * heads) == 0):
        return False
    return qkv.shape[-1] // (3 * heads) in (64, 128) and qkv.shape[0] * heads <= 65535


def _args(qkv: Tensor, heads: int, scale: float, out: Tensor, lse: Tensor) -> _capi.LsFmha:
    B, L, C3 = qkv.shape
    HD = C3 // 3
    D = HD // heads
    base = qkv.data_ptr()
    return _capi.LsFmha(B, heads, L, D, scale, 0, C3, C3, C3, out.stride(1), base, base + 4 * HD, base + 8 * HD, out.data_ptr(),
                        lse.data_ptr())


class _AttentionPacked(torch.autograd.Function):
    @staticmethod
    def forward(ctx, qkv: Tensor, heads: int, scale: float) -> Tensor:
        qkv = qkv.contiguous()
        B, L, C3 = qkv.shape
        out = torch.empty((B, L, C3 // 3), dtype=torch.float32, device=qkv.device)
        lse = torch.empty((B, heads, L), dtype=torch.float32, device=qkv.device)
        a = _args(qkv, heads, scale, out, lse)
        with torch.cuda.device(qkv.device):
            _capi.check(_capi.load().ls_fmha_forward(C.byref(a), torch.cuda.current_stream().cuda_stream), "ls_fmha_forward")
        _capi.KERNEL_LAUNCHES[0] += 1
        ctx.save_for_backward(qkv, out, lse)
        ctx.cfg = (heads, scale)
        return out

    @staticmethod
    def backward(ctx, g: Tensor):
        qkv, out, lse = ctx.saved_tensors
        heads, scale = ctx.cfg
        g = g.contiguous()
        B, L, C3 = qkv.shape
        HD = C3 // 3
        dqkv = torch.empty_like(qkv)
        delta = torch.empty_like(lse)
        a = _args(qkv, heads, scale, out, lse)
        base = dqkv.data_ptr()
        with torch.cuda.device(qkv.device):
            _capi.check(_capi.load().ls_fmha_backward(C.byref(a), g.data_ptr(), base, base + 4 * HD, base + 8 * HD, delta.data_ptr(),
                                                      torch.cuda.current_stream().cuda_stream), "ls_fmha_backward")
        _capi.KERNEL_LAUNCHES[0] += 3
        return dqkv, None, None


def attention_packed(qkv: Tensor, heads: int, scale: float) -> Tensor:
    if not qkv.is_cuda:
        raise RuntimeError("attention_packed needs CUDA tensors: latentsplat_b200 has no CPU fallback")
    return _AttentionPacked.apply(qkv, heads, scale)
