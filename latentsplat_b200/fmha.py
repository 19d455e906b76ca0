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
from .gemm import gemm_tf32

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
        _capi.FLOPS["fmha"] += 4.0 * B * heads * L * L * (C3 // 3 // heads)          # S = Q K^T and O = P V
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
        _capi.FLOPS["fmha"] += 10.0 * B * heads * L * L * (HD // heads)              # algorithmic: dV, dP, dQ, dK + S recompute
        return dqkv, None, None


def attention_packed(qkv: Tensor, heads: int, scale: float) -> Tensor:
    if not qkv.is_cuda:
        raise RuntimeError("attention_packed needs CUDA tensors: latentsplat_b200 has no CPU fallback")
    return _AttentionPacked.apply(qkv, heads, scale)


# ---- wide single-head attention (VAE mid block): scores through our GEMM, softmax in place ------------------------------
def wide_supported(q: Tensor, k: Tensor, v: Tensor) -> bool:
    ok = ENABLED and q.is_cuda and q.dim() == 3 and q.shape == k.shape == v.shape
    ok = ok and all(t.dtype == torch.float32 and t.is_contiguous() for t in (q, k, v))
    return bool(ok and q.shape[1] % 4 == 0 and q.shape[1] <= 4096 and q.shape[2] % 4 == 0)


def _softmax_rows(x: Tensor, scale: float) -> None:
    with torch.cuda.device(x.device):
        _capi.check(_capi.load().ls_softmax_rows_forward(x.data_ptr(), x.shape[0], x.shape[1], x.stride(0), scale,
                                                         torch.cuda.current_stream().cuda_stream), "ls_softmax_rows_forward")
    _capi.KERNEL_LAUNCHES[0] += 1


class _AttentionWide(torch.autograd.Function):
    """softmax(q k^T scale) v for (b, L, c) single-head tensors with a wide head (c = 512 in the VAE): six ls_gemm_tf32 calls
    per batch element (S = Q K^T, O = P V; dV = P^T dO, dP = dO V^T, dQ = dS K, dK = dS^T Q -- operands read K-major or
    MN-major in place) around the in-place row-softmax kernels.  P (b, L, L) is kept for the backward."""

    @staticmethod
    def forward(ctx, q: Tensor, k: Tensor, v: Tensor, scale: float) -> Tensor:
        b, L, c = q.shape
        p = torch.empty((b, L, L), dtype=torch.float32, device=q.device)
        out = torch.empty_like(q)
        for i in range(b):
            gemm_tf32(q[i], k[i], M=L, N=L, K=c, out=p[i], split_k=1)
            _softmax_rows(p[i], scale)
            gemm_tf32(p[i], v[i], M=L, N=c, K=L, b_mn=True, out=out[i], split_k=1)
        ctx.save_for_backward(q, k, v, p)
        ctx.scale = scale
        return out

    @staticmethod
    def backward(ctx, g: Tensor):
        q, k, v, p = ctx.saved_tensors
        b, L, c = q.shape
        g = g.contiguous()
        dq, dk, dv = torch.empty_like(q), torch.empty_like(k), torch.empty_like(v)
        ds = torch.empty((L, L), dtype=torch.float32, device=q.device)
        lib = _capi.load()
        for i in range(b):
            gemm_tf32(p[i], g[i], M=L, N=c, K=L, a_mn=True, b_mn=True, out=dv[i], split_k=1)       # dV = P^T dO
            gemm_tf32(g[i], v[i], M=L, N=L, K=c, out=ds, split_k=1)                                # dP = dO V^T
            with torch.cuda.device(q.device):
                _capi.check(lib.ls_softmax_rows_backward(p[i].data_ptr(), ds.data_ptr(), L, L, L, ctx.scale,
                                                         torch.cuda.current_stream().cuda_stream), "ls_softmax_rows_backward")
            _capi.KERNEL_LAUNCHES[0] += 1
            gemm_tf32(ds, k[i], M=L, N=c, K=L, b_mn=True, out=dq[i], split_k=1)                    # dQ = dS K
            gemm_tf32(ds, q[i], M=L, N=c, K=L, a_mn=True, b_mn=True, out=dk[i], split_k=1)         # dK = dS^T Q
        return dq, dk, dv, None


def attention_wide(q: Tensor, k: Tensor, v: Tensor, scale: float) -> Tensor:
    if not q.is_cuda:
        raise RuntimeError("attention_wide needs CUDA tensors: latentsplat_b200 has no CPU fallback")
    return _AttentionWide.apply(q, k, v, scale)
