"""The abstract interfaces of the render path in one place.

The reference spreads one small abstract base per package (`src/model/encoder/encoder.py:13-41`,
`encoder/backbone/backbone.py:12-36`, `decoder/decoder.py:11-54`, `autoencoder/autoencoder.py:12-69`,
`discriminator/discriminator.py:10-33`); their public surface -- constructor arguments, method names, properties --
is what makes this package a drop-in, so it is restated here, on a shared `Configured` base.  The per-package modules
(`encoder/encoder.py`, ...) re-export these names so that imports written against the reference layout keep working.
"""
from __future__ import annotations

from abc import ABC, abstractmethod
from dataclasses import dataclass
from fractions import Fraction
from typing import Callable, Generic, Literal, Optional, TypeVar

from torch import Tensor, nn

from .diagonal_gaussian_distribution import DiagonalGaussianDistribution
from .types import Gaussians, VariationalGaussians

CfgT = TypeVar("CfgT")
DepthRenderingMode = Literal["depth", "log", "disparity", "relative_disparity"]


class Configured(nn.Module, ABC, Generic[CfgT]):
    """An nn.Module that keeps the (dataclass) config it was built from as `.cfg`."""
    cfg: CfgT

    def __init__(self, cfg: CfgT) -> None:
        super().__init__()
        self.cfg = cfg


# ---- encoder side ------------------------------------------------------------------------------------------------
class Backbone(Configured[CfgT]):
    """(batch, d_in, height, width) images -> (batch, d_out, height * scale_factor, width * scale_factor) features."""

    def __init__(self, cfg: CfgT, d_in: int, d_out: int, scale_factor: Fraction) -> None:
        super().__init__(cfg)
        self.d_in, self.d_out, self.scale_factor = d_in, d_out, scale_factor

    @abstractmethod
    def forward(self, x: Tensor) -> Tensor: ...


class Encoder(Configured[CfgT]):
    """context views -> variational Gaussians; `variational` doubles the feature channels (mean, log-variance)."""
    variational: bool

    def __init__(self, cfg: CfgT, variational: bool) -> None:
        super().__init__(cfg)
        self.variational = variational

    @abstractmethod
    def forward(self, context: dict, global_step: int, features: Optional[Tensor] = None, deterministic: bool = False,
                visualization_dump: Optional[dict] = None) -> VariationalGaussians: ...

    def get_data_shim(self) -> Callable[[dict], dict]:
        return lambda batch: batch                      # identity unless an encoder needs patch-aligned / bounded inputs

    @property
    @abstractmethod
    def last_layer_weights(self) -> Optional[Tensor]: ...


# ---- decoder side ------------------------------------------------------------------------------------------------
@dataclass
class DecoderOutput:
    color: Optional[Tensor]                                        # (batch, view, 3, h, w)
    feature_posterior: Optional[DiagonalGaussianDistribution]      # over (batch, view, c, h, w)
    mask: Tensor                                                   # (batch, view, h, w)
    depth: Tensor                                                  # (batch, view, h, w)


class Decoder(Configured[CfgT]):
    @abstractmethod
    def forward(self, gaussians: Gaussians, extrinsics: Tensor, intrinsics: Tensor, near: Tensor, far: Tensor,
                image_shape: tuple[int, int], depth_mode: Optional[DepthRenderingMode] = None, return_colors: bool = True,
                return_features: bool = True) -> DecoderOutput: ...

    @property
    @abstractmethod
    def last_layer_weights(self) -> Optional[Tensor]: ...


class Autoencoder(Configured[CfgT]):
    @abstractmethod
    def encode(self, images: Tensor) -> DiagonalGaussianDistribution: ...

    @abstractmethod
    def decode(self, z: Tensor, skip_z: Optional[Tensor] = None) -> Tensor: ...

    @property
    @abstractmethod
    def downscale_factor(self) -> int: ...

    @property
    @abstractmethod
    def d_latent(self) -> int: ...

    @property
    @abstractmethod
    def last_layer_weights(self) -> Optional[Tensor]: ...

    @property
    @abstractmethod
    def expects_skip(self) -> bool: ...

    @property
    @abstractmethod
    def expects_skip_extra(self) -> bool: ...


class Discriminator(Configured[CfgT]):
    """(batch, in_dim, h, w) -> (batch, 1, ~h / downscale_factor, ~w / downscale_factor) patch logits."""

    @abstractmethod
    def forward(self, input: Tensor) -> Tensor: ...

    @property
    @abstractmethod
    def downscale_factor(self) -> int: ...
