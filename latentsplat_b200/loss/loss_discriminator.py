"""Discriminator loss on fake and real logits, hinge or vanilla (/root/reference/src/loss/loss_discriminator.py:16-70).
Returns a {"fake", "real"} dict, each weighted with half the configured weight."""
from dataclasses import dataclass
from typing import Literal

import torch
import torch.nn.functional as F
from torch import Tensor

from .loss import Loss, LossCfg, LossOutput, LossValue


@dataclass
class LossDiscriminatorCfg(LossCfg):
    name: Literal["discriminator"] = "discriminator"
    loss: Literal["hinge", "vanilla"] = "hinge"


class LossDiscriminator(Loss):
    def __init__(self, cfg: LossDiscriminatorCfg) -> None:
        super().__init__(cfg)
        self.loss = {"hinge": self.hinge_loss, "vanilla": self.vanilla_loss}[cfg.loss]

    @staticmethod
    def hinge_loss(logits: Tensor) -> Tensor:
        return F.relu(1.0 + logits).mean()

    @staticmethod
    def vanilla_loss(logits: Tensor) -> Tensor:
        return F.softplus(logits).mean()

    def unweighted_loss(self, prediction, gt=None):
        return self.loss(prediction.logits_fake), self.loss(-prediction.logits_real)     # real logits enter negated

    def forward(self, prediction, gt=None, global_step: int = 0) -> LossOutput:
        if self.is_active(global_step):
            fake, real = self.unweighted_loss(prediction, gt)
        else:
            fake = real = torch.zeros((), dtype=torch.float32, device=prediction.device)
        half = self.cfg.weight / 2
        return {"fake": LossValue(fake, half * fake), "real": LossValue(real, half * real)}
