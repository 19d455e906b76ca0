"""TF32 tcgen05 GEMM (libls_raster.so::ls_gemm_tf32, include/ls_gemm.h) and the `Linear` layer built on it.

`Linear` is a drop-in `nn.Linear` (same parameter names, so reference checkpoints load) whose CUDA forward,
input-gradient and weight-gradient GEMMs run on our sm_100a kernel with bias (and ReLU) fused into the epilogue;
it replaces the fp32 SIMT cuBLAS GEMMs the reference gets from torch's default matmul precision.  On CPU tensors
(host-logic tests) it is plain `F.linear`.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import torch
import torch.nn.functional as F
from torch import Tensor, nn

from . import _capi

ACT = {"none": _capi.ACT_NONE, "relu": _capi.ACT_RELU, "gelu": _capi.ACT_GELU, "silu": _capi.ACT_SILU}
NUM_SMS = 148
enabled = True          # set False to route CUDA linears through cuBLAS (A/B comparisons)


def gemm_tf32(A: Tensor, B: Tensor, *, M: int, N: int, K: int, a_mn: bool = False, b_mn: bool = False,
              bias: Optional[Tensor] = None, act: str = "none", out: Optional[Tensor] = None, split_k: int = 0,
              accumulate: bool = False, residual: Optional[Tensor] = None, pre_out: Optional[Tensor] = None) -> Tensor:
    """out[M,N] (+)= act(op(A)[M,K] @ op(B)[N,K]^T (+ bias)) (+ residual).   A: (M,K) row-major, or (K,M) if a_mn; B likewise.
    `split_k=0` picks a split that fills the 148 SMs when the output has few tiles (weight gradients).
    `residual` (M,N): added after the activation (a block's skip connection); `pre_out` (M,N): receives the pre-activation."""
    if not (A.is_cuda and B.is_cuda):
        raise RuntimeError("gemm_tf32 needs CUDA tensors: latentsplat_b200 has no CPU fallback")
    assert A.dtype == torch.float32 and B.dtype == torch.float32 and A.dim() == 2 and B.dim() == 2
    assert A.stride(1) == 1 and B.stride(1) == 1, "operands must be row-major with unit inner stride"
    assert tuple(A.shape) == ((K, M) if a_mn else (M, K)) and tuple(B.shape) == ((K, N) if b_mn else (N, K))
    if out is None:
        out = torch.empty((M, N), dtype=torch.float32, device=A.device)
    if split_k == 0:
        tiles = ((M + 127) // 128) * ((N + 127) // 128)
        nkb = (K + 31) // 32
        split_k = 1 if tiles >= NUM_SMS else max(1, min(nkb // 4, (2 * NUM_SMS + tiles - 1) // tiles))
        if act != "none" or (bias is not None and split_k > 1) or residual is not None or pre_out is not None:
            split_k = 1
    if residual is not None:
        assert residual.shape == (M, N) and residual.stride(1) == 1 and residual.dtype == torch.float32
    if pre_out is not None:
        assert pre_out.shape == (M, N) and pre_out.stride() == out.stride()
    args = _capi.LsGemmArgs(M, N, K, int(a_mn), int(b_mn), ACT[act], int(split_k), int(accumulate), A.stride(0),
                            B.stride(0), out.stride(0), A.data_ptr(), B.data_ptr(), out.data_ptr(),
                            None if bias is None else bias.data_ptr(),
                            None if residual is None else residual.data_ptr(), 0 if residual is None else residual.stride(0),
                            None if pre_out is None else pre_out.data_ptr())
    with torch.cuda.device(A.device):
        _capi.check(_capi.load().ls_gemm_tf32(C.byref(args), torch.cuda.current_stream().cuda_stream), "ls_gemm_tf32")
    _capi.KERNEL_LAUNCHES[0] += 1
    _capi.FLOPS["gemm"] += 2.0 * M * N * K
    return out


def col_sum(x2d: Tensor) -> Tensor:
    """x2d.sum(dim=0) for a row-major CUDA fp32 matrix on libls_raster.so::ls_col_sum (bias gradients)."""
    if not (x2d.is_cuda and x2d.dtype == torch.float32 and x2d.dim() == 2 and x2d.stride(1) == 1 and x2d.stride(0) % 4 == 0
            and x2d.data_ptr() % 16 == 0 and x2d.shape[0] > 0):
        return x2d.sum(dim=0)
    out = torch.zeros(x2d.shape[1], dtype=torch.float32, device=x2d.device)
    with torch.cuda.device(x2d.device):
        _capi.check(_capi.load().ls_col_sum(x2d.data_ptr(), out.data_ptr(), x2d.shape[0], x2d.shape[1], x2d.stride(0),
                                            torch.cuda.current_stream().cuda_stream), "ls_col_sum")
    _capi.KERNEL_LAUNCHES[0] += 1
    return out


def _aligned(t: Tensor) -> bool:
    return t.data_ptr() % 16 == 0 and t.stride(0) % 4 == 0


class _LinearFn(torch.autograd.Function):
    """y = act(x W^T + b) (+ residual): bias, ReLU / GELU / SiLU and the skip connection run in the GEMM epilogue; the smooth
    activations keep their pre-activation (second epilogue output) for the backward pass."""

    @staticmethod
    def forward(ctx, x2d: Tensor, weight: Tensor, bias: Optional[Tensor], act: str, residual: Optional[Tensor]):
        M, N = x2d.shape[0], weight.shape[0]
        pre = torch.empty((M, N), dtype=torch.float32, device=x2d.device) if act in ("gelu", "silu") else None
        y = gemm_tf32(x2d, weight, M=M, N=N, K=x2d.shape[1], bias=bias, act=act, residual=residual, pre_out=pre)
        ctx.act = act
        ctx.has_bias = bias is not None
        # relu: the output's sign IS the mask -- unless a residual was added on top of it
        if act == "relu" and residual is not None:
            raise RuntimeError("relu + residual is not wired (no such layer on the path)")
        ctx.save_for_backward(x2d, weight, y if act == "relu" else pre)
        return y

    @staticmethod
    def backward(ctx, gy: Tensor):
        x2d, weight, saved = ctx.saved_tensors
        gy = gy.contiguous()
        g_res = gy if ctx.needs_input_grad[4] else None              # the skip connection passes the gradient through
        if ctx.act == "relu":
            gy = gy * (saved > 0)
        elif ctx.act in ("gelu", "silu"):
            from .conv import act_backward                            # elementwise dy * act'(pre) on our kernel
            gy = act_backward(gy, saved, ctx.act)
        M, K = x2d.shape
        N = weight.shape[0]
        gx = gw = gb = None
        if ctx.needs_input_grad[0]:      # dX = dY W     : A = dY (K-major over N), B = W read MN-major
            gx = gemm_tf32(gy, weight, M=M, N=K, K=N, b_mn=True)
        if ctx.needs_input_grad[1]:      # dW = dY^T X   : both operands MN-major, reduce over M (split-K)
            gw = gemm_tf32(gy, x2d, M=N, N=K, K=M, a_mn=True, b_mn=True)
        if ctx.has_bias and ctx.needs_input_grad[2]:
            gb = col_sum(gy)
        return gx, gw, gb, None, g_res


class _GroupedLinearFn(torch.autograd.Function):
    """y[g] = x[g] @ w[g]^T + b[g] for g groups with their own weights (x (G, M, K), w (G, N, K), b (G, N)): one autograd
    node, G GEMM launches writing straight into slices of one output / one input-gradient tensor (slicing at the autograd
    level would add a zero-fill + accumulate pass over the whole activation per group)."""

    @staticmethod
    def forward(ctx, x: Tensor, weight: Tensor, bias: Tensor):
        G, M, K = x.shape
        N = weight.shape[1]
        y = torch.empty((G, M, N), dtype=torch.float32, device=x.device)
        for g in range(G):
            gemm_tf32(x[g], weight[g], M=M, N=N, K=K, bias=bias[g], out=y[g])
        ctx.save_for_backward(x, weight)
        return y

    @staticmethod
    def backward(ctx, gy: Tensor):
        x, weight = ctx.saved_tensors
        G, M, K = x.shape
        N = weight.shape[1]
        gy = gy.contiguous()
        gx = torch.empty_like(x) if ctx.needs_input_grad[0] else None
        gw = torch.empty_like(weight) if ctx.needs_input_grad[1] else None
        for g in range(G):
            if gx is not None:
                gemm_tf32(gy[g], weight[g], M=M, N=K, K=N, b_mn=True, out=gx[g])
            if gw is not None:
                gemm_tf32(gy[g], x[g], M=N, N=K, K=M, a_mn=True, b_mn=True, out=gw[g])
        gb = torch.stack([col_sum(gy[g]) for g in range(G)]) if ctx.needs_input_grad[2] else None
        return gx, gw, gb


def grouped_linear(x: Tensor, weight: Tensor, bias: Tensor) -> Tensor:
    """x (G, M, K), weight (G, N, K), bias (G, N) -> (G, M, N); CUDA fp32 on the tcgen05 GEMM, else a batched matmul."""
    if (x.is_cuda and enabled and x.dtype == torch.float32 and x.shape[2] % 4 == 0 and weight.shape[1] % 4 == 0
            and x.shape[1] > 0):
        return _GroupedLinearFn.apply(x.contiguous(), weight.contiguous(), bias.contiguous())
    return torch.baddbmm(bias[:, None], x, weight.transpose(1, 2))


def linear(x: Tensor, weight: Tensor, bias: Optional[Tensor] = None, act: str = "none",
           residual: Optional[Tensor] = None) -> Tensor:
    """act(F.linear(x, weight, bias)) (+ residual) on the tcgen05 GEMM when possible (CUDA fp32, TMA-compatible strides);
    act in "none" | "relu" | "gelu" | "silu".  `residual` has the output's shape."""
    N = weight.shape[0]
    if x.is_cuda and enabled and x.dtype == torch.float32 and weight.dtype == torch.float32:
        x2d = x.reshape(-1, x.shape[-1])
        if not x2d.is_contiguous():
            x2d = x2d.contiguous()
        K = x2d.shape[1]
        r2d = None
        if residual is not None:
            r2d = residual.reshape(-1, N)
            if not r2d.is_contiguous():
                r2d = r2d.contiguous()
        if K % 4 == 0 and N % 4 == 0 and _aligned(x2d) and _aligned(weight) and x2d.shape[0] > 0:
            return _LinearFn.apply(x2d, weight, bias, act, r2d).reshape(*x.shape[:-1], N)
    y = F.linear(x, weight, bias)
    y = F.relu(y) if act == "relu" else F.gelu(y) if act == "gelu" else F.silu(y) if act == "silu" else y
    return y if residual is None else y + residual


class Linear(nn.Linear):
    """nn.Linear whose CUDA path is our tcgen05 TF32 GEMM (parameters: `weight`, `bias`, as nn.Linear).  `act` fuses the
    following activation, `residual=` at call time the block's skip connection, into the GEMM epilogue."""

    def __init__(self, *args, act: str = "none", **kwargs):
        super().__init__(*args, **kwargs)
        assert act in ACT
        self.act = act

    def forward(self, input: Tensor, residual: Optional[Tensor] = None) -> Tensor:
        return linear(input, self.weight, self.bias, self.act, residual)
