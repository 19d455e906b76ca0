"""Depth smoothness regulariser (/root/reference/src/loss/loss_depth.py:13-56): first (optionally second) differences of the
log-depth normalised to [near, far], optionally damped where the ground-truth image has an edge."""
from dataclasses import dataclass
from typing import Literal, Optional

import torch
from torch import Tensor

from .loss import Loss, LossCfg


@dataclass
class LossDepthCfg(LossCfg):
    name: Literal["depth"] = "depth"
    sigma_image: Optional[float] = None
    use_second_derivative: bool = False


class LossDepth(Loss):
    def unweighted_loss(self, prediction, gt) -> Tensor:
        lo, hi = gt.near[..., None, None].log(), gt.far[..., None, None].log()
        # prediction.depth is compared with LOG near / far (the reference clamps the raw depth map between them, :27-30)
        d = (prediction.depth.minimum(hi).maximum(lo) - lo) / (hi - lo)
        dx, dy = d.diff(dim=-1), d.diff(dim=-2)
        second = self.cfg.use_second_derivative
        if second:
            dx, dy = dx.diff(dim=-1), dy.diff(dim=-2)
        if self.cfg.sigma_image is not None:
            cx = gt.image.diff(dim=-1).amax(dim=2)              # strongest channel difference, (b, v, h, w-1)
            cy = gt.image.diff(dim=-2).amax(dim=2)
            if second:
                cx = torch.maximum(cx[..., :, 1:], cx[..., :, :-1])
                cy = torch.maximum(cy[..., 1:, :], cy[..., :-1, :])
            dx = dx * torch.exp(-self.cfg.sigma_image * cx)
            dy = dy * torch.exp(-self.cfg.sigma_image * cy)
        return dx.abs().mean() + dy.abs().mean()
