"""Pre-norm transformer stack with its two small building blocks.

Parameter tree of /root/reference/src/model/transformer/{transformer.py:32-71, pre_norm.py:28-35, feed_forward.py:28-40}:
`layers[i] = ModuleList([PreNorm(Attention), PreNorm(feed_forward_layer)])` with `PreNorm.{norm, fn}` and
`FeedForward.net = Sequential(Linear, GELU, Dropout, Linear, Dropout)`;  x <- attn(x, z) + x;  x <- ff(x) + x.
In cross-attention only x is normalised, z (keys/values) is used raw.  LayerNorm and the Linears run on our sm_100a
kernels on CUDA (latentsplat_b200.norm / .gemm)."""
from __future__ import annotations

from typing import Callable, Optional

from torch import Tensor, nn

from latentsplat_b200.gemm import Linear
from latentsplat_b200.norm import LayerNorm

from .attention import Attention


class FeedForward(nn.Module):
    """Token MLP dim -> hidden_dim -> dim with GELU."""

    def __init__(self, dim: int, hidden_dim: int, dropout: float = 0.0) -> None:
        super().__init__()
        # reference: Linear, GELU, Dropout, Linear, Dropout.  The GELU runs in the first Linear's GEMM epilogue (an Identity
        # keeps the Sequential indices net.0 / net.3 of the checkpoint), the block's skip connection in the second one's.
        stages = [Linear(dim, hidden_dim, act="gelu"), nn.Identity(), nn.Dropout(dropout), Linear(hidden_dim, dim),
                  nn.Dropout(dropout)]
        self.net = nn.Sequential(*stages)
        self.fused_residual = dropout == 0.0          # (x + Dropout(y) != Dropout(x + y) when dropout is active)

    def forward(self, x: Tensor, residual: Optional[Tensor] = None) -> Tensor:
        if residual is None or not self.fused_residual:
            y = self.net(x)
            return y if residual is None else y + residual
        h = self.net[2](self.net[1](self.net[0](x)))
        return self.net[4](self.net[3](h, residual=residual))


class PreNorm(nn.Module):
    """fn(LayerNorm(x), **kwargs): `norm` normalises the token dimension, extra keyword arguments go to `fn` untouched."""

    def __init__(self, dim: int, fn: nn.Module) -> None:
        super().__init__()
        self.norm = LayerNorm(dim)
        self.fn = fn

    def forward(self, x: Tensor, **kwargs) -> Tensor:
        return self.fn(self.norm(x), **kwargs)


class Transformer(nn.Module):
    def __init__(self, dim: int, depth: int, heads: int, dim_head: int, mlp_dim: int, dropout: float = 0.0,
                 selfatt: bool = True, kv_dim: Optional[int] = None,
                 feed_forward_layer: Callable[..., nn.Module] = FeedForward) -> None:
        super().__init__()
        def block() -> nn.ModuleList:
            attention = Attention(dim, heads=heads, dim_head=dim_head, dropout=dropout, selfatt=selfatt, kv_dim=kv_dim)
            return nn.ModuleList([PreNorm(dim, attention), PreNorm(dim, feed_forward_layer(dim, mlp_dim, dropout=dropout))])
        self.layers = nn.ModuleList([block() for _ in range(depth)])

    def forward(self, x: Tensor, z: Optional[Tensor] = None, **kwargs) -> Tensor:
        for attn, ff in self.layers:
            x = attn(x, z=z) + x
            x = ff(x, residual=x) if isinstance(ff.fn, FeedForward) else ff(x, **kwargs) + x
        return x
