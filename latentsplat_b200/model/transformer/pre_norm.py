"""LayerNorm-then-function wrapper (/root/reference/src/model/transformer/pre_norm.py:28-35)."""
from torch import nn

from latentsplat_b200.norm import LayerNorm  # nn.LayerNorm on our warp-per-row kernel on CUDA


class PreNorm(nn.Module):
    def __init__(self, dim, fn):
        super().__init__()
        self.norm = LayerNorm(dim)
        self.fn = fn

    def forward(self, x, **kwargs):
        return self.fn(self.norm(x), **kwargs)
