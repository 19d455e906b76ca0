"""Generator (non-saturating hinge) term: minus the mean PatchGAN logit of the rendered image
(/root/reference/src/loss/loss_generator.py:15-30)."""
from dataclasses import dataclass
from typing import Literal

from torch import Tensor

from .loss import Loss, LossCfg


@dataclass
class LossGeneratorCfg(LossCfg):
    name: Literal["generator"] = "generator"


class LossGenerator(Loss):
    def unweighted_loss(self, prediction, gt=None) -> Tensor:
        return -prediction.logits_fake.mean()
