"""Epipolar cross-attention transformer (/root/reference/src/model/encoder/epipolar/epipolar_transformer.py:18-183).

Per ray of every context view: one query token (the pixel's feature), `num_samples` key/value tokens sampled
along the epipolar line in the other view(s) with a positional encoding of their triangulated depth, then a
convolutional feed-forward with image self-attention; optional 4x down/up-scaling around it.
The reference reads `num_context_views` from a global hydra config (:47); here it is a constructor argument
(default 2, the value of every shipped experiment).
"""
from __future__ import annotations

from dataclasses import dataclass
from functools import partial
from typing import Optional

from torch import Tensor, nn

from ....geometry.epipolar_lines import get_depth
from ...encodings.positional_encoding import PositionalEncoding
from ...transformer.transformer import Transformer
from .conversions import depth_to_relative_disparity
from .epipolar_sampler import EpipolarSampler, EpipolarSampling
from .image_self_attention import ImageSelfAttention, ImageSelfAttentionCfg
from latentsplat_b200.gemm import Linear, linear  # nn.Linear / F.linear with tcgen05 TF32 GEMMs on CUDA
from latentsplat_b200 import epipolar_gather as fused_gather
from latentsplat_b200.conv import Conv2d, ConvTranspose2d  # tcgen05 implicit-GEMM convolutions (NHWC) with fused bias + activation


@dataclass
class EpipolarTransformerCfg:
    self_attention: ImageSelfAttentionCfg
    num_octaves: int
    num_layers: int
    num_heads: int
    num_samples: int
    d_dot: int
    d_mlp: int
    downscale: int


class ConvFeedForward(nn.Module):
    """layers(self_attention(x) + x) on the (b v, c, h, w) view of the token sequence (:155-183)."""

    def __init__(self, self_attention_cfg: ImageSelfAttentionCfg, d_in: int, d_hidden: int, dropout: float) -> None:
        super().__init__()
        # reference: Conv, GELU, Dropout, Conv, Dropout (:164-170).  The GELU runs in the first convolution's epilogue; an
        # Identity keeps the Sequential indices (checkpoint keys layers.0 / layers.3).
        self.layers = nn.Sequential(Conv2d(d_in, d_hidden, 7, 1, 3, act="gelu"), nn.Identity(), nn.Dropout(dropout),
                                    Conv2d(d_hidden, d_in, 7, 1, 3), nn.Dropout(dropout))
        self.self_attention = ImageSelfAttention(self_attention_cfg, d_in, d_in)

    def forward(self, x: Tensor, b: int, v: int, h: int, w: int) -> Tensor:
        c = x.shape[-1]
        x = x.reshape(b * v, h, w, c).permute(0, 3, 1, 2)          # (b v h w) () c -> (b v) c h w
        x = self.layers(self.self_attention(x) + x)
        return x.permute(0, 2, 3, 1).reshape(b * v * h * w, 1, c)


class EpipolarTransformer(nn.Module):
    def __init__(self, cfg: EpipolarTransformerCfg, d_in: int, num_context_views: int = 2) -> None:
        super().__init__()
        self.cfg = cfg
        self.epipolar_sampler = EpipolarSampler(num_context_views, cfg.num_samples)
        if cfg.num_octaves > 0:
            pe = PositionalEncoding(cfg.num_octaves)
            self.depth_encoding = nn.Sequential(pe, Linear(pe.d_out(1), d_in))
        self.transformer = Transformer(d_in, cfg.num_layers, cfg.num_heads, cfg.d_dot, cfg.d_mlp, selfatt=False,
                                       kv_dim=d_in, feed_forward_layer=partial(ConvFeedForward, cfg.self_attention))
        if cfg.downscale > 1:
            self.downscaler = Conv2d(d_in, d_in, cfg.downscale, cfg.downscale)
            self.upscaler = ConvTranspose2d(d_in, d_in, cfg.downscale, cfg.downscale)
            self.upscale_refinement = nn.Sequential(Conv2d(d_in, d_in * 2, 7, 1, 3, act="gelu"), nn.Identity(),
                                                    Conv2d(d_in * 2, d_in, 7, 1, 3))      # Conv, GELU (fused), Conv (:70-74)
        else:
            self.downscaler = self.upscaler = self.upscale_refinement = None

    def _downscale(self, features: Tensor, repeats: int) -> Tensor:
        """downscaler(repeat(features, repeats)) for (bv, c, h, w) features.  When every `repeats x repeats` block of the
        full-resolution map is one replicated pixel (BackboneDino's "repeat" up-sampling, backbone_dino.py:72-84) and the
        stride-`downscale` filter windows do not straddle blocks, the strided convolution sees a constant window:
            conv(x_rep)[y, x] = (sum_{r,s} W[:, :, r, s]) x[y', x'] + b
        i.e. a 1x1 Linear with the tap-summed weight on the coarse grid, replicated repeats/downscale times -- exact, and the
        256x256x128 replicated tensor (268 MB at the bench shape) is never formed nor convolved (8.6 GMAC -> 0.13)."""
        d = self.cfg.downscale
        conv = self.downscaler
        if repeats > 1 and repeats % d == 0:
            w_sum = conv.weight.sum(dim=(2, 3))                                        # (c_out, c_in); autograd spreads the gradient over the taps
            y = linear(features.permute(0, 2, 3, 1), w_sum, conv.bias)                  # (bv, h, w, c_out) token-major
            k = repeats // d
            if k > 1:
                bv, hc, wc, c = y.shape
                y = y[:, :, None, :, None, :].expand(bv, hc, k, wc, k, c).reshape(bv, hc * k, wc * k, c)
            return y.permute(0, 3, 1, 2)                                                # NCHW view of NHWC memory (channels_last)
        if repeats > 1:
            features = features.repeat_interleave(repeats, dim=2).repeat_interleave(repeats, dim=3)
        return conv(features)

    def forward(self, features: Tensor, extrinsics: Tensor, intrinsics: Tensor, near: Tensor, far: Tensor, repeats: int = 1):
        """`repeats` > 1: `features` is a coarse map whose pixels each stand for repeats x repeats identical full-resolution
        pixels (the encoder passes the backbone's un-replicated token grid)."""
        b, v, c, h, w = features.shape
        h, w = h * repeats, w * repeats
        if self.downscaler is not None:
            features = self._downscale(features.flatten(0, 1), repeats).unflatten(0, (b, v))
        elif repeats > 1:
            features = features.repeat_interleave(repeats, dim=3).repeat_interleave(repeats, dim=4)
        hd, wd = h // self.cfg.downscale, w // self.cfg.downscale

        encoding = self.depth_encoding[1] if self.cfg.num_octaves > 0 else None
        fused = fused_gather.supported(features, c, self.cfg.num_samples, None if encoding is None else encoding.weight)
        sampling: EpipolarSampling = self.epipolar_sampler(features, extrinsics, intrinsics, near, far, gather=not fused)
        depths = None
        if self.cfg.num_octaves > 0:
            collect = self.epipolar_sampler.collect
            depths = get_depth(sampling.origins[:, :, None, :, None], sampling.directions[:, :, None, :, None],
                               sampling.xy_sample, collect(extrinsics)[:, :, :, None, None],
                               collect(intrinsics)[:, :, :, None, None])
            # clip: context views may be extremely close together or oriented the same way (:113-116)
            depths = depths.maximum(near[..., None, None, None]).minimum(far[..., None, None, None])
            depths = depth_to_relative_disparity(depths, near[..., None, None, None], far[..., None, None, None])
            if not fused:
                q = sampling.features + self.depth_encoding(depths[..., None])
        elif not fused:
            q = sampling.features

        # NB the reference's names: `kv` are the per-pixel query tokens x, `q` the sampled key/value tokens z.
        channels_last = features.permute(0, 1, 3, 4, 2).contiguous()          # (b, v, hd, wd, c): query tokens AND gather source
        kv = channels_last.reshape(b * v * hd * wd, 1, c)
        assert v == 2, "the reference's rearrange 'b v () r s c' admits exactly one other view"
        if fused:
            # one sm_100a kernel: bilinear gather on the epipolar line in the other view * overlap mask + depth encoding
            z = fused_gather.epipolar_gather(
                channels_last.reshape(b * v, hd, wd, c), sampling.xy_sample.reshape(b * v * hd * wd, -1, 2),
                None if depths is None else depths.expand(sampling.xy_sample.shape[:-1]).reshape(b * v * hd * wd, -1),
                self.epipolar_sampler.image_index(b, hd * wd), sampling.valid.reshape(-1).to(features.dtype),
                None if encoding is None else encoding.weight, None if encoding is None else encoding.bias)
        else:
            z = q.reshape(b * v * hd * wd, -1, c)                 # b v () r s c -> (b v r) s c
        features = self.transformer(kv, z, b=b, v=v, h=hd, w=wd)
        features = features.reshape(b, v, hd, wd, c).permute(0, 1, 4, 2, 3)

        if self.upscaler is not None:
            x = self.upscaler(features.flatten(0, 1))
            x = self.upscale_refinement(x) + x
            features = x.unflatten(0, (b, v))
        return features, sampling
