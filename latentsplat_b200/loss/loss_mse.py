"""Mean squared image error (/root/reference/src/loss/loss_mse.py:11-23)."""
from dataclasses import dataclass
from typing import Literal

from torch import Tensor

from .loss import Loss, LossCfg


@dataclass
class LossMseCfg(LossCfg):
    name: Literal["mse"] = "mse"


class LossMse(Loss):
    def unweighted_loss(self, prediction, gt) -> Tensor:
        return (prediction.image - gt.image).square().mean()
