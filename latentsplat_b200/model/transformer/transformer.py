"""Pre-norm transformer stack (/root/reference/src/model/transformer/transformer.py:32-71):
layers[i] = ModuleList([PreNorm(Attention), PreNorm(feed_forward_layer)]); x <- attn(x, z)+x; x <- ff(x)+x.
Note that in cross-attention only x is normalised, z (keys/values) is used raw."""
from torch import nn

from .attention import Attention
from .feed_forward import FeedForward
from .pre_norm import PreNorm


class Transformer(nn.Module):
    def __init__(self, dim, depth, heads, dim_head, mlp_dim, dropout=0.0, selfatt=True, kv_dim=None,
                 feed_forward_layer=FeedForward):
        super().__init__()
        self.layers = nn.ModuleList([
            nn.ModuleList([
                PreNorm(dim, Attention(dim, heads=heads, dim_head=dim_head, dropout=dropout, selfatt=selfatt,
                                       kv_dim=kv_dim)),
                PreNorm(dim, feed_forward_layer(dim, mlp_dim, dropout=dropout)),
            ]) for _ in range(depth)])

    def forward(self, x, z=None, **kwargs):
        for attn, ff in self.layers:
            x = attn(x, z=z) + x
            x = ff(x, **kwargs) + x
        return x
