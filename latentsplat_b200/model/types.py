"""Data carriers crossing the encoder/decoder boundary.

Field names and shapes follow /root/reference/src/model/types.py:9-58.
"""
from __future__ import annotations

from dataclasses import dataclass, fields
from typing import Literal, Optional

from torch import Tensor

from .diagonal_gaussian_distribution import DiagonalGaussianDistribution


@dataclass
class Gaussians:
    means: Tensor                              # (batch, gaussian, 3)
    covariances: Tensor                        # (batch, gaussian, 3, 3)
    opacities: Tensor                          # (batch, gaussian)
    color_harmonics: Optional[Tensor] = None   # (batch, gaussian, 3, d_color_sh)
    feature_harmonics: Optional[Tensor] = None # (batch, gaussian, channels, d_feature_sh)


@dataclass
class VariationalGaussians(Gaussians):
    feature_harmonics: Optional[DiagonalGaussianDistribution] = None

    def _to_gaussians(self, feature_harmonics: Tensor) -> Gaussians:
        return Gaussians(self.means, self.covariances, self.opacities, self.color_harmonics, feature_harmonics)

    def flatten(self) -> Gaussians:
        return self._to_gaussians(self.feature_harmonics.params)

    def mode(self) -> Gaussians:
        return self._to_gaussians(self.feature_harmonics.mode())

    def sample(self) -> Gaussians:
        return self._to_gaussians(self.feature_harmonics.sample())


@dataclass
class Prediction:
    image: Optional[Tensor] = None
    posterior: Optional[DiagonalGaussianDistribution] = None
    depth: Optional[Tensor] = None
    logits_fake: Optional[Tensor] = None
    logits_real: Optional[Tensor] = None

    @property
    def device(self):
        for f in fields(self):
            val = getattr(self, f.name)
            if val is not None:
                return val.device


@dataclass
class GroundTruth:
    image: Optional[Tensor] = None
    near: Optional[Tensor] = None
    far: Optional[Tensor] = None


VariationalMode = Literal["none", "gaussians", "latents"]
