"""KL divergence of the rendered latent posterior to the unit Gaussian (/root/reference/src/loss/loss_kl.py:11-22)."""
from dataclasses import dataclass
from typing import Literal

from torch import Tensor

from .loss import Loss, LossCfg


@dataclass
class LossKlCfg(LossCfg):
    name: Literal["kl"] = "kl"


class LossKl(Loss):
    def unweighted_loss(self, prediction, gt=None) -> Tensor:
        return prediction.posterior.kl().mean()
