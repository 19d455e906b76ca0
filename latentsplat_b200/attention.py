"""Fused single-query attention (libls_raster.so::ls_sq_attention_*, include/ls_gemm.h) as an autograd function."""
from __future__ import annotations

import torch
from torch import Tensor

from . import _capi

ABSORB = True      # use the weight-absorbed cross-attention when applicable (set False for A/B comparisons)
ENABLED = True     # set False to route the epipolar cross-attention through the explicit torch sequence (bench comparator)


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
    return (ENABLED and q.is_cuda and q.dtype == torch.float32 and kv.dtype == torch.float32 and q.dim() == 2 and kv.dim() == 3
            and q.shape[1] == heads * 128 and kv.shape[2] == 2 * heads * 128 and kv.shape[1] <= 32)


# ------------------------------------------------------------------------------------------------------------------
# Weight-absorbed epipolar cross-attention: never forms kv = to_kv(z)
# ------------------------------------------------------------------------------------------------------------------
class _AbsorbedCrossAttention(torch.autograd.Function):
    """out = concat_h  W_v,h ( sum_j softmax_j( (W_k,h^T q_h) . z_j * scale ) z_j ),  exactly
    softmax(q k^T scale) v with k, v = to_kv(z).chunk(2) when to_kv has no bias (attention.py:45-46, 60-68)."""

    @staticmethod
    def forward(ctx, q: Tensor, z: Tensor, w_kv: Tensor, heads: int, scale: float):
        from .gemm import gemm_tf32
        R, HD = q.shape
        D = HD // heads
        S, Dz = z.shape[1], z.shape[2]
        q, z, w_kv = q.contiguous(), z.contiguous(), w_kv.contiguous()
        wk, wv = w_kv[:HD], w_kv[HD:]                                   # (H*D, Dz) each, row = h*D + d
        qt = torch.empty((R, heads * Dz), dtype=torch.float32, device=q.device)
        for h in range(heads):                                          # qt_h = q_h @ W_k,h   (W_k,h stored [d][c])
            gemm_tf32(q[:, h * D:(h + 1) * D], wk[h * D:(h + 1) * D], M=R, N=Dz, K=D, b_mn=True, out=qt[:, h * Dz:(h + 1) * Dz],
                      split_k=1)
        zbar = torch.empty_like(qt)
        p = torch.empty((R, heads, S), dtype=torch.float32, device=q.device)
        lib = _capi.load()
        with torch.cuda.device(q.device):
            _capi.check(lib.ls_absorbed_attention_forward(qt.data_ptr(), z.data_ptr(), zbar.data_ptr(), p.data_ptr(), R, heads,
                                                          S, Dz, scale, torch.cuda.current_stream().cuda_stream),
                        "ls_absorbed_attention_forward")
        _capi.KERNEL_LAUNCHES[0] += 1
        out = torch.empty((R, HD), dtype=torch.float32, device=q.device)
        for h in range(heads):                                          # out_h = zbar_h @ W_v,h^T
            gemm_tf32(zbar[:, h * Dz:(h + 1) * Dz], wv[h * D:(h + 1) * D], M=R, N=D, K=Dz, out=out[:, h * D:(h + 1) * D], split_k=1)
        ctx.save_for_backward(q, z, w_kv, qt, zbar, p)
        ctx.cfg = (heads, scale)
        return out

    @staticmethod
    def backward(ctx, dout: Tensor):
        from .gemm import gemm_tf32
        q, z, w_kv, qt, zbar, p = ctx.saved_tensors
        heads, scale = ctx.cfg
        R, HD = q.shape
        D = HD // heads
        S, Dz = z.shape[1], z.shape[2]
        dout = dout.contiguous()
        wk, wv = w_kv[:HD], w_kv[HD:]
        dw = torch.empty_like(w_kv)
        dzbar = torch.empty_like(zbar)
        for h in range(heads):
            sl_d, sl_z = slice(h * D, (h + 1) * D), slice(h * Dz, (h + 1) * Dz)
            gemm_tf32(dout[:, sl_d], wv[sl_d], M=R, N=Dz, K=D, b_mn=True, out=dzbar[:, sl_z], split_k=1)      # dzbar_h = dout_h W_v,h
            gemm_tf32(dout[:, sl_d], zbar[:, sl_z], M=D, N=Dz, K=R, a_mn=True, b_mn=True, out=dw[HD + h * D:HD + (h + 1) * D])  # dW_v,h
        dqt, dz = torch.empty_like(qt), torch.empty_like(z)
        lib = _capi.load()
        with torch.cuda.device(q.device):
            _capi.check(lib.ls_absorbed_attention_backward(qt.data_ptr(), z.data_ptr(), p.data_ptr(), dzbar.data_ptr(),
                                                           dqt.data_ptr(), dz.data_ptr(), R, heads, S, Dz, scale,
                                                           torch.cuda.current_stream().cuda_stream), "ls_absorbed_attention_backward")
        _capi.KERNEL_LAUNCHES[0] += 1
        dq = torch.empty_like(q)
        for h in range(heads):
            sl_d, sl_z = slice(h * D, (h + 1) * D), slice(h * Dz, (h + 1) * Dz)
            gemm_tf32(dqt[:, sl_z], wk[sl_d], M=R, N=D, K=Dz, out=dq[:, sl_d], split_k=1)                      # dq_h = dqt_h W_k,h^T
            gemm_tf32(q[:, sl_d], dqt[:, sl_z], M=D, N=Dz, K=R, a_mn=True, b_mn=True, out=dw[h * D:(h + 1) * D])  # dW_k,h
        return dq, dz, dw, None, None


def absorbed_cross_attention(q: Tensor, z: Tensor, w_kv: Tensor, heads: int, scale: float) -> Tensor:
    """q (R, H*D) projected queries, z (R, S<=32, 128) raw key/value samples, w_kv (2*H*D, 128) the bias-free to_kv weight."""
    return _AbsorbedCrossAttention.apply(q, z, w_kv, heads, scale)


def absorbed_supported(q: Tensor, z: Tensor, w_kv: Tensor, heads: int) -> bool:
    return (ENABLED and q.is_cuda and q.dtype == z.dtype == w_kv.dtype == torch.float32 and q.dim() == 2 and z.dim() == 3
            and z.shape[2] == 128 and z.shape[1] <= 32 and heads <= 8 and q.shape[1] % heads == 0
            and (q.shape[1] // heads) % 4 == 0 and w_kv.shape == (2 * q.shape[1], 128))
