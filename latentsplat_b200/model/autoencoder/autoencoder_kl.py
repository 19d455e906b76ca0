"""KL autoencoder with high-resolution skip injection (/root/reference/src/model/autoencoder/autoencoder_kl.py:25-200).

`decode(z, skip_z)`: post_quant_conv -> conv_in -> mid block -> before every up block add
`skip_convs[i](bilinear(skip_z -> z.size, align_corners=True))` (1x1 convs, zero-initialised) -> up blocks ->
GroupNorm, SiLU, conv_out -> (x + 1) / 2   (:93-124, :168-180).  As in the reference, one more skip conv is
created than the decoder loop uses (:66-74).
"""
from __future__ import annotations

import os
from dataclasses import dataclass, field
from typing import Literal, Optional

import torch
from torch import Tensor, nn
from torch.nn.functional import interpolate

from ..diagonal_gaussian_distribution import DiagonalGaussianDistribution
from .autoencoder import Autoencoder
from .vae_kl import AutoencoderKLModel
from latentsplat_b200.conv import Conv2d  # tcgen05 implicit-GEMM convolutions (NHWC) with fused bias + activation

PRETRAINED_AUTOENCODER_PATH = "pretrained/autoencoder"     # /root/reference/src/constants.py:1


@dataclass
class AutoencoderKLCfg:
    name: Literal["kl"]
    model: Literal["kl_f8", "kl_f16", "kl_f32"]
    down_block_types: list[str]
    up_block_types: list[str]
    block_out_channels: list[int]
    layers_per_block: int
    latent_channels: int
    skip_connections: bool = False
    skip_extra: bool = True
    skip_zero: bool = True
    pretrained: bool = True


def zero_module(module: nn.Module) -> nn.Module:
    for p in module.parameters():
        nn.init.zeros_(p)
    return module


class AutoencoderKL(Autoencoder[AutoencoderKLCfg]):
    def __init__(self, cfg: AutoencoderKLCfg, d_in: int = 3, d_skip_extra: int = 0, sample_size: int = 32) -> None:
        super().__init__(cfg)
        assert all(t == "DownEncoderBlock2D" for t in cfg.down_block_types)
        assert all(t == "UpDecoderBlock2D" for t in cfg.up_block_types)
        self.model = AutoencoderKLModel(d_in, d_in, tuple(cfg.block_out_channels), cfg.layers_per_block,
                                        cfg.latent_channels)
        if cfg.pretrained:
            state_dict = torch.load(os.path.join(PRETRAINED_AUTOENCODER_PATH, cfg.model + ".pt"), map_location="cpu")
            self.model.load_state_dict(state_dict)
        if cfg.skip_connections:
            self.d_skip = self.d_latent + (d_skip_extra if cfg.skip_extra else 0)
            mk = lambda d_out: (zero_module if cfg.skip_zero else (lambda m: m))(Conv2d(self.d_skip, d_out, kernel_size=1))
            self.skip_convs = nn.ModuleList([mk(cfg.block_out_channels[-1])] +
                                            [mk(d) for d in reversed(cfg.block_out_channels)])

    def encode(self, images: Tensor) -> DiagonalGaussianDistribution:
        """Images in [0,1]; keeps the leading batch dimensions."""
        batch_dims = images.shape[:-3]
        mean, logvar = self.model.encode_moments((2 * images - 1).flatten(0, -4))
        return DiagonalGaussianDistribution(mean=mean.reshape(*batch_dims, *mean.shape[1:]),
                                            logvar=logvar.reshape(*batch_dims, *logvar.shape[1:]))

    def _decoder_forward(self, z: Tensor, skip_z: Optional[Tensor] = None) -> Tensor:
        decoder = self.model.decoder
        z = decoder.mid_block(decoder.conv_in(z))
        for i, up_block in enumerate(decoder.up_blocks):
            if self.cfg.skip_connections:
                z = z + self.skip_convs[i](interpolate(skip_z, size=z.shape[-2:], mode="bilinear", align_corners=True))
            z = up_block(z)
        return decoder.conv_out(decoder.conv_norm_out(z))      # conv_act (SiLU) is fused into conv_norm_out (vae_kl.py)

    def decode(self, z: Tensor, skip_z: Optional[Tensor] = None) -> Tensor:
        batch_dims = z.shape[:-3]
        z = z.flatten(0, -4)
        if skip_z is not None:
            skip_z = skip_z.flatten(0, -4)
        sample = self._decoder_forward(self.model.post_quant_conv(z), skip_z)
        sample = (sample + 1) / 2
        return sample.reshape(*batch_dims, *sample.shape[1:])

    @property
    def downscale_factor(self) -> int:
        return 2 ** (len(self.cfg.block_out_channels) - 1)

    @property
    def d_latent(self) -> int:
        return self.cfg.latent_channels

    @property
    def last_layer_weights(self) -> Tensor:
        return self.model.decoder.conv_out.weight

    @property
    def expects_skip(self) -> bool:
        return self.cfg.skip_connections

    @property
    def expects_skip_extra(self) -> bool:
        return self.cfg.skip_extra
