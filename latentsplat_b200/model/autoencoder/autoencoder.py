"""Autoencoder base class (/root/reference/src/model/autoencoder/autoencoder.py:12-69)."""
from abc import ABC, abstractmethod
from typing import Generic, Optional, TypeVar

from torch import Tensor, nn

from ..diagonal_gaussian_distribution import DiagonalGaussianDistribution

T = TypeVar("T")


class Autoencoder(nn.Module, ABC, Generic[T]):
    cfg: T

    def __init__(self, cfg: T) -> None:
        super().__init__()
        self.cfg = cfg

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
