"""Discriminator base class (/root/reference/src/model/discriminator/discriminator.py:10-33)."""
from abc import ABC, abstractmethod
from typing import Generic, TypeVar

from torch import Tensor, nn

T = TypeVar("T")


class Discriminator(nn.Module, ABC, Generic[T]):
    cfg: T

    def __init__(self, cfg: T) -> None:
        super().__init__()
        self.cfg = cfg

    @abstractmethod
    def forward(self, input: Tensor) -> Tensor:
        """(batch, in_dim, h, w) -> (batch, 1, ~h / downscale, ~w / downscale) patch logits."""

    @property
    @abstractmethod
    def downscale_factor(self) -> int: ...
