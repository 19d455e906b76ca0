"""A named group of losses: NLL-type terms, an optional generator term with the adaptive GAN weight, an optional discriminator
term (/root/reference/src/loss/loss_group.py:14-118; the adaptive weight is the VQGAN rule the reference uses at :34-45 from
model_wrapper.py:958-984: ||d nll / d W_last|| / (||d g / d W_last|| + 1e-4), clamped to [0, 1], detached).

The two `torch.autograd.grad` calls re-traverse the graph down to the decoder's last layer only (retain_graph=True); with our
convolution kernels that is one extra wgrad launch per call."""
from __future__ import annotations

from typing import Dict, List, Optional, Tuple, Union

import torch
from torch import Tensor, nn

from .loss import Loss, LossValue
from .loss_discriminator import LossDiscriminator
from .loss_generator import LossGenerator


def flatten_nested(d: dict, prefix: Optional[str] = None) -> Dict[str, LossValue]:
    """{"a": {"b": v}} -> {"prefix/a/b": v} (/root/reference/src/misc/dict_utils.py:6-17)."""
    out = {}
    for key, value in d.items():
        path = key if prefix is None else f"{prefix}/{key}"
        if isinstance(value, dict):
            out.update(flatten_nested(value, path))
        else:
            out[path] = value
    return out


class LossGroup(nn.Module):
    def __init__(self, name: str, nll_losses: Optional[List[Union[Loss, "LossGroup"]]] = None,
                 generator_loss: Optional[LossGenerator] = None, discriminator_loss: Optional[LossDiscriminator] = None) -> None:
        super().__init__()
        self.name = name
        self.nll_losses = nn.ModuleList(nll_losses or [])
        self.generator_loss = generator_loss
        self.discriminator_loss = discriminator_loss

    @staticmethod
    def get_adaptive_weight(nll_loss: Tensor, g_loss: Tensor, last_layer_weights: Tensor) -> Tensor:
        (nll_grad,) = torch.autograd.grad(nll_loss, last_layer_weights, retain_graph=True)
        (g_grad,) = torch.autograd.grad(g_loss, last_layer_weights, retain_graph=True)
        return (torch.linalg.norm(nll_grad) / (torch.linalg.norm(g_grad) + 1e-4)).clamp(0.0, 1.0).detach()

    def forward(self, prediction, gt=None, global_step: int = 0):
        return self.forward_generator(prediction, gt, global_step)

    def forward_generator(self, prediction, gt=None, global_step: int = 0,
                          last_layer_weights: Optional[Tensor] = None) -> Tuple[Union[Tensor, int], Dict[str, LossValue]]:
        losses = flatten_nested({l.name: l(prediction, gt, global_step) for l in self.nll_losses}, self.name)
        total = sum(v.weighted for v in losses.values())
        if self.is_generator_loss_active(global_step):
            g = self.generator_loss(prediction, gt, global_step)
            g.weighted = self.get_adaptive_weight(total, g.unweighted, last_layer_weights) * g.weighted
            total = total + g.weighted
            losses[f"{self.name}/{self.generator_loss.name}"] = g
        return total, losses

    def forward_discriminator(self, prediction, gt, global_step: int) -> Tuple[Union[Tensor, int], Dict[str, LossValue]]:
        losses = flatten_nested(self.discriminator_loss(prediction, gt, global_step), self.name)
        return sum(v.weighted for v in losses.values()), losses

    @property
    def has_generator_loss(self) -> bool:
        return self.generator_loss is not None

    @property
    def has_discriminator_loss(self) -> bool:
        return self.discriminator_loss is not None

    def is_generator_loss_active(self, global_step: int) -> bool:
        return self.has_generator_loss and self.generator_loss.is_active(global_step)

    def is_discriminator_loss_active(self, global_step: int) -> bool:
        return self.has_discriminator_loss and self.discriminator_loss.is_active(global_step)

    def is_active(self, global_step: int) -> bool:
        return (any(l.is_active(global_step) for l in self.nll_losses) or self.is_generator_loss_active(global_step)
                or self.is_discriminator_loss_active(global_step))
