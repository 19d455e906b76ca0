"""Loss heads of the training step (/root/reference/src/loss/__init__.py:14-57): the same registry, config dataclasses and
`get_loss_group(name, cfg)` factory; see SURVEY.md section 8(f) rank 1."""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Optional, Union

from .loss import Loss, LossCfg, LossValue
from .loss_depth import LossDepth, LossDepthCfg
from .loss_discriminator import LossDiscriminator, LossDiscriminatorCfg
from .loss_generator import LossGenerator, LossGeneratorCfg
from .loss_group import LossGroup
from .loss_kl import LossKl, LossKlCfg
from .loss_l1 import LossL1, LossL1Cfg
from .loss_lpips import LossLpips, LossLpipsCfg, LpipsVgg
from .loss_mse import LossMse, LossMseCfg

LOSSES = {"depth": LossDepth, "kl": LossKl, "l1": LossL1, "lpips": LossLpips, "mse": LossMse}

NLLLossCfg = Union[LossDepthCfg, LossKlCfg, LossL1Cfg, LossLpipsCfg, LossMseCfg]


@dataclass
class LossGroupCfg:
    nll: Optional[List[NLLLossCfg]] = None
    generator: Optional[LossGeneratorCfg] = None
    discriminator: Optional[LossDiscriminatorCfg] = None


def get_loss_group(name: str, group_cfg: Optional[LossGroupCfg] = None) -> LossGroup:
    if group_cfg is None:
        return LossGroup(name)
    nll, shared_lpips = [], None
    for cfg in group_cfg.nll or []:
        if cfg.name == "lpips":                # one VGG for every LPIPS term of the group
            loss = LossLpips(cfg, lpips=shared_lpips)
            shared_lpips = loss.lpips
        else:
            loss = LOSSES[cfg.name](cfg)
        nll.append(loss)
    return LossGroup(name, nll,
                     generator_loss=None if group_cfg.generator is None else LossGenerator(group_cfg.generator),
                     discriminator_loss=None if group_cfg.discriminator is None else LossDiscriminator(group_cfg.discriminator))


__all__ = ["Loss", "LossCfg", "LossValue", "LossGroup", "LossGroupCfg", "get_loss_group", "LOSSES", "LossDepth", "LossDepthCfg",
           "LossDiscriminator", "LossDiscriminatorCfg", "LossGenerator", "LossGeneratorCfg", "LossKl", "LossKlCfg", "LossL1",
           "LossL1Cfg", "LossLpips", "LossLpipsCfg", "LpipsVgg", "LossMse", "LossMseCfg"]
