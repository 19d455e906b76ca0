"""Patch-token self-attention over a feature map
(/root/reference/src/model/encoder/epipolar/image_self_attention.py:13-79)."""
from __future__ import annotations

from dataclasses import dataclass

from torch import Tensor, nn

from ....geometry.projection import sample_image_grid
from ...encodings.positional_encoding import PositionalEncoding
from ...transformer.transformer import Transformer
from latentsplat_b200.gemm import Linear  # nn.Linear with tcgen05 TF32 GEMMs on CUDA
from latentsplat_b200.conv import Conv2d, ConvTranspose2d  # tcgen05 implicit-GEMM convolutions (NHWC) with fused bias + activation


@dataclass
class ImageSelfAttentionCfg:
    patch_size: int
    num_octaves: int
    num_layers: int
    num_heads: int
    d_token: int
    d_dot: int
    d_mlp: int


class ImageSelfAttention(nn.Module):
    def __init__(self, cfg: ImageSelfAttentionCfg, d_in: int, d_out: int):
        super().__init__()
        pe = PositionalEncoding(cfg.num_octaves)
        self.positional_encoding = nn.Sequential(pe, Linear(pe.d_out(2), cfg.d_token))
        self.patch_embedder = nn.Sequential(Conv2d(d_in, cfg.d_token, cfg.patch_size, cfg.patch_size, act="relu"), nn.Identity())
        self.transformer = Transformer(cfg.d_token, cfg.num_layers, cfg.num_heads, cfg.d_dot, cfg.d_mlp)
        self.resampler = ConvTranspose2d(cfg.d_token, d_out, cfg.patch_size, cfg.patch_size)

    def forward(self, image: Tensor) -> Tensor:
        tokens = self.patch_embedder(image)
        _, _, nh, nw = tokens.shape
        xy, _ = sample_image_grid((nh, nw), device=image.device)
        tokens = tokens + self.positional_encoding(xy).permute(2, 0, 1)
        tokens = self.transformer(tokens.flatten(2).transpose(1, 2))            # b c nh nw -> b (nh nw) c
        tokens = tokens.transpose(1, 2).unflatten(2, (nh, nw))
        return self.resampler(tokens)
