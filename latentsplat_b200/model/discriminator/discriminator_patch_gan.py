"""PatchGAN discriminator as in pix2pix (/root/reference/src/model/discriminator/discriminator_patch_gan.py:14-114):
conv k4 s2 (+bias) -> LeakyReLU; (n_layers-1) x [conv k4 s2, BatchNorm, LeakyReLU]; conv k4 s1, BatchNorm,
LeakyReLU; conv k4 s1 -> 1 (+bias).  All layers live in `self.main` (checkpoint key prefix `main.N`)."""
from __future__ import annotations

import os
from dataclasses import dataclass
from typing import Literal

import torch
from torch import Tensor, nn

from .discriminator import Discriminator
from latentsplat_b200.conv import Conv2d  # tcgen05 implicit-GEMM convolutions (NHWC) with fused bias + activation

PRETRAINED_DISCRIMINATOR_PATH = "pretrained/discriminator"


@dataclass
class DiscriminatorPatchGanCfg:
    name: Literal["patch_gan"]
    model: Literal["kl_f8", "kl_f16", "kl_f32"]
    base_dim: int = 64
    max_dim_mult: int = 8
    n_layers: int = 3
    downscale_factor: int = 2
    kernel_size: int = 4
    padding: int = 1
    leaky_relu_neg_slope: float = 0.2
    pretrained: bool = True


class DiscriminatorPatchGan(Discriminator[DiscriminatorPatchGanCfg]):
    def __init__(self, cfg: DiscriminatorPatchGanCfg, d_in: int = 3):
        super().__init__(cfg)
        c = self.cfg
        act = lambda: nn.LeakyReLU(negative_slope=c.leaky_relu_neg_slope, inplace=True)
        conv = lambda i, o, s, bias, a="none": Conv2d(i, o, kernel_size=c.kernel_size, stride=s, padding=c.padding, bias=bias, act=a)
        # the first LeakyReLU follows its convolution directly: fused into the epilogue when the slope is the kernel's 0.2
        # (an Identity keeps the Sequential indices main.N of the checkpoint)
        fuse = c.leaky_relu_neg_slope == 0.2
        layers = [conv(d_in, c.base_dim, c.downscale_factor, True, "lrelu" if fuse else "none"), nn.Identity() if fuse else act()]
        mult = 1
        for n in range(1, c.n_layers):
            prev, mult = mult, min(c.downscale_factor ** n, c.max_dim_mult)
            layers += [conv(c.base_dim * prev, c.base_dim * mult, c.downscale_factor, False),
                       nn.BatchNorm2d(c.base_dim * mult), act()]
        prev, mult = mult, min(c.downscale_factor ** c.n_layers, c.max_dim_mult)
        layers += [conv(c.base_dim * prev, c.base_dim * mult, 1, False), nn.BatchNorm2d(c.base_dim * mult), act(),
                   conv(c.base_dim * mult, 1, 1, True)]
        self.main = nn.Sequential(*layers)
        if c.pretrained:
            self.load_state_dict(torch.load(os.path.join(PRETRAINED_DISCRIMINATOR_PATH, c.model + ".pt"),
                                            map_location="cpu"))

    def forward(self, input: Tensor) -> Tensor:
        return self.main(input)

    @property
    def downscale_factor(self) -> int:
        return self.cfg.downscale_factor ** self.cfg.n_layers
