"""Token MLP (/root/reference/src/model/transformer/feed_forward.py:28-40)."""
from torch import nn
from latentsplat_b200.gemm import Linear  # nn.Linear with tcgen05 TF32 GEMMs on CUDA


class FeedForward(nn.Module):
    def __init__(self, dim, hidden_dim, dropout=0.0):
        super().__init__()
        self.net = nn.Sequential(Linear(dim, hidden_dim), nn.GELU(), nn.Dropout(dropout),
                                 Linear(hidden_dim, dim), nn.Dropout(dropout))

    def forward(self, x):
        return self.net(x)
