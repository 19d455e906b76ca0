"""Loss base class and value carrier (interface of /root/reference/src/loss/loss.py:12-63).

A loss maps (Prediction, GroundTruth) to a scalar; `forward` returns the unweighted and the weighted value and is a constant 0
before `apply_after_step`.  Everything here is device-agnostic torch: on CUDA the reductions run as the few elementwise /
reduction kernels of torch inside the step's CUDA graph (they are a rounding error next to the encoder / decoder kernels).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, Optional, Union

import torch
from torch import Tensor, nn

from ..model.types import GroundTruth, Prediction


@dataclass
class LossCfg:
    name: str
    weight: Union[float, int] = 1
    apply_after_step: int = 0


@dataclass
class LossValue:
    unweighted: Tensor
    weighted: Tensor


LossOutput = Union[LossValue, Dict[str, "LossOutput"]]      # nested dicts are flattened by LossGroup


class Loss(nn.Module):
    def __init__(self, cfg: LossCfg) -> None:
        super().__init__()
        self.cfg = cfg
        self.name = cfg.name

    def unweighted_loss(self, prediction: Prediction, gt: Optional[GroundTruth] = None) -> Tensor:
        raise NotImplementedError

    def is_active(self, global_step: int) -> bool:
        return global_step >= self.cfg.apply_after_step

    def forward(self, prediction: Prediction, gt: Optional[GroundTruth] = None, global_step: int = 0) -> LossOutput:
        if self.is_active(global_step):
            value = self.unweighted_loss(prediction, gt)
        else:
            value = torch.zeros((), dtype=torch.float32, device=prediction.device)
        return LossValue(value, self.cfg.weight * value)
