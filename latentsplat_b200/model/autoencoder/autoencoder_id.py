"""Identity autoencoder (/root/reference/src/model/autoencoder/autoencoder_id.py:11-59)."""
from dataclasses import dataclass
from typing import Literal, Optional

from torch import Tensor

from ..diagonal_gaussian_distribution import DiagonalGaussianDistribution
from .autoencoder import Autoencoder


@dataclass
class AutoencoderIdCfg:
    name: Literal["id"]
    skip_connections: bool = False


class AutoencoderId(Autoencoder[AutoencoderIdCfg]):
    def __init__(self, cfg: AutoencoderIdCfg, d_in: int = 3, d_skip_extra: int = 0, sample_size: int = 32) -> None:
        super().__init__(cfg)
        self.d_in = d_in

    def encode(self, images: Tensor) -> DiagonalGaussianDistribution:
        return DiagonalGaussianDistribution(images)

    def decode(self, z: Tensor, skip_z: Optional[Tensor] = None) -> Tensor:
        return z

    downscale_factor = property(lambda self: 1)
    d_latent = property(lambda self: self.d_in)
    last_layer_weights = property(lambda self: None)
    expects_skip = property(lambda self: False)
    expects_skip_extra = property(lambda self: False)
