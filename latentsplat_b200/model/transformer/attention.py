"""Multi-head (self or cross) attention with the parameter tree of
/root/reference/src/model/transformer/attention.py:30-70: `to_qkv` (self) or `to_q`/`to_kv` (cross), all
without bias, `to_out` = Sequential(Linear(+bias), Dropout), an `attend` Softmax module (the reference's
visualiser hooks it, encoder_visualizer_epipolar.py:55-58)."""
from __future__ import annotations

import torch
from torch import nn

from latentsplat_b200 import attention as fused
from latentsplat_b200 import fmha
from latentsplat_b200.gemm import Linear  # nn.Linear with tcgen05 TF32 GEMMs on CUDA


class Attention(nn.Module):
    def __init__(self, dim, heads=8, dim_head=64, dropout=0.0, selfatt=True, kv_dim=None):
        super().__init__()
        inner_dim = dim_head * heads
        project_out = not (heads == 1 and dim_head == dim)
        self.heads = heads
        self.scale = dim_head ** -0.5
        self.attend = nn.Softmax(dim=-1)
        if selfatt:
            self.to_qkv = Linear(dim, inner_dim * 3, bias=False)
        else:
            self.to_q = Linear(dim, inner_dim, bias=False)
            self.to_kv = Linear(kv_dim, inner_dim * 2, bias=False)
        self.to_out = nn.Sequential(Linear(inner_dim, dim), nn.Dropout(dropout)) if project_out else nn.Identity()

    def forward(self, x, z=None):
        if z is None:
            qkv = self.to_qkv(x)
            if not self.attend._forward_hooks and fmha.supported(qkv, self.heads):
                # self-attention (ImageSelfAttention: 4 heads x 128): tcgen05 flash-attention core on the packed projection
                return self.to_out(fmha.attention_packed(qkv, self.heads, self.scale))
            q, k, v = qkv.chunk(3, dim=-1)
        else:
            q = self.to_q(x)
            # Epipolar cross-attention: one query per ray.  (1) weight-absorbed form: to_kv is folded into the query and
            # the output projection of the attended samples, the 4.3 GB kv tensor is never formed; (2) else the fused
            # single-query kernel on kv; (3) else (hooks on `attend` for the visualiser, CPU, odd shapes) the explicit path.
            plain = x.shape[1] == 1 and not self.attend._forward_hooks and self.to_kv.bias is None
            if plain and fused.ABSORB and fused.absorbed_supported(q[:, 0], z, self.to_kv.weight, self.heads):
                return self.to_out(fused.absorbed_cross_attention(q[:, 0], z, self.to_kv.weight, self.heads, self.scale)[:, None])
            kv = self.to_kv(z)
            if plain and fused.supported(q[:, 0], kv, self.heads):
                return self.to_out(fused.single_query_attention(q[:, 0], kv, self.heads, self.scale)[:, None])
            k, v = kv.chunk(2, dim=-1)
        split = lambda t: t.unflatten(-1, (self.heads, -1)).transpose(1, 2)      # b n (h d) -> b h n d
        q, k, v = split(q), split(k), split(v)
        dots = torch.matmul(q, k.transpose(-1, -2)) * self.scale
        out = torch.matmul(self.attend(dots), v)
        return self.to_out(out.transpose(1, 2).flatten(-2))                      # b h n d -> b n (h d)
