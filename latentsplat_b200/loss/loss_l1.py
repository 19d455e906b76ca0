"""Mean absolute image error (/root/reference/src/loss/loss_l1.py:11-23)."""
from dataclasses import dataclass
from typing import Literal

from torch import Tensor

from .loss import Loss, LossCfg


@dataclass
class LossL1Cfg(LossCfg):
    name: Literal["l1"] = "l1"


class LossL1(Loss):
    def unweighted_loss(self, prediction, gt) -> Tensor:
        return (prediction.image - gt.image).abs().mean()
