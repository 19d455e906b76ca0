"""Backbone factory (/root/reference/src/model/encoder/backbone/__init__.py).  Only DINO is on the hot path
(config/model/encoder/backbone/dino.yaml); resnet / vit / ensemble are out of scope."""
from fractions import Fraction

from .backbone import Backbone
from .backbone_dino import BackboneDino, BackboneDinoCfg

BACKBONES = {"dino": BackboneDino}
BackboneCfg = BackboneDinoCfg


def get_backbone(cfg: BackboneCfg, d_in: int, d_out: int, scale_factor: Fraction) -> Backbone:
    if isinstance(cfg, list) or cfg.name not in BACKBONES:
        raise NotImplementedError(f"backbone {cfg!r}: only the DINO backbone of the shipped experiments is built")
    return BACKBONES[cfg.name](cfg, d_in, d_out, scale_factor)
