"""Encoder factory (/root/reference/src/model/encoder/__init__.py:16-33); the visualiser slot is always None."""
from fractions import Fraction
from typing import Optional

from .encoder import Encoder
from .encoder_epipolar import EncoderEpipolar, EncoderEpipolarCfg

ENCODERS = {"epipolar": EncoderEpipolar}
EncoderCfg = EncoderEpipolarCfg


def get_encoder(cfg: EncoderCfg, d_in: int, n_feature_channels: int, scale_factor: Fraction,
                variational: bool = False) -> tuple[Encoder, Optional[object]]:
    return ENCODERS[cfg.name](cfg, d_in, n_feature_channels, scale_factor, variational), None
