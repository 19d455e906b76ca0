"""Encoder base class (/root/reference/src/model/encoder/encoder.py:13-41)."""
from abc import ABC, abstractmethod
from typing import Generic, Optional, TypeVar

from torch import Tensor, nn

from ..types import VariationalGaussians

T = TypeVar("T")


class Encoder(nn.Module, ABC, Generic[T]):
    cfg: T
    variational: bool

    def __init__(self, cfg: T, variational: bool) -> None:
        super().__init__()
        self.cfg = cfg
        self.variational = variational

    @abstractmethod
    def forward(self, context: dict, global_step: int, features: Optional[Tensor] = None,
                deterministic: bool = False, visualization_dump: Optional[dict] = None) -> VariationalGaussians:
        ...

    def get_data_shim(self):
        """The default shim doesn't modify the batch."""
        return lambda x: x

    @property
    @abstractmethod
    def last_layer_weights(self) -> Optional[Tensor]:
        ...
