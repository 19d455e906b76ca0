"""Sync-free training step: the optimisation part of `ModelWrapper.training_step`
(/root/reference/src/model/model_wrapper.py:424-489) for the render path -- generator pass (render pipeline -> loss groups with the
adaptive GAN weight -> backward -> gradient-norm clip -> Adam) followed by the discriminator pass on the detached prediction
(hinge / vanilla on fake and real logits -> backward -> clip -> Adam) -- without a single host synchronisation:

  * the reference's `if not loss.isnan().any(): ... step()` guard (:440, :473) is evaluated ON THE DEVICE: a non-finite loss
    zeroes that pass's gradients and multiplies its learning rate by 0 for this step, so the parameters do not move (the Adam
    moments see a zero gradient and the step counter advances -- the one documented difference to "skip");
  * gradients live in one flat buffer per optimiser (parallel.FlatGradients): clipping is two kernels, the data-parallel
    all-reduce one collective; the ACTIVE parameter set is fixed at construction (no `find_unused_parameters` traversal);
  * no `.item()` prints: losses come back as device scalars, the caller reads them when it wants to.

Both passes have static shapes and no host-side control flow, so each (or the whole step) can be captured by
runtime.GraphedStep, as bench.py does for the generator pass.  SURVEY.md section 8(f) rank 2.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Callable, Dict, Iterable, Optional

import torch
from torch import Tensor, nn

from .loss import LossGroup
from .model.types import GroundTruth, Prediction
from .parallel import FlatGradients


@dataclass
class OptimizerCfg:
    """`optimizer.generator` / `optimizer.discriminator` of config/main.yaml:79-90."""
    lr: float = 1.5e-5
    gradient_clip_val: Optional[float] = 0.5
    betas: tuple = (0.9, 0.999)


class GuardedAdam:
    """Adam over a FlatGradients buffer with a device-side learning rate: `step(ok)` with ok = 0 leaves the parameters unchanged."""

    def __init__(self, params: Iterable[nn.Parameter], cfg: OptimizerCfg):
        self.cfg = cfg
        self.grads = FlatGradients(params)
        dev = self.grads.flat.device
        self.lr = torch.tensor(cfg.lr, device=dev)                # a tensor lr is read on the device (capturable)
        fused = dev.type == "cuda"
        self.opt = torch.optim.Adam(self.grads.params, lr=self.lr, betas=cfg.betas, fused=fused, capturable=fused)

    def zero(self) -> None:
        self.grads.zero()

    def step(self, ok: Tensor, reduce: bool = True) -> Tensor:
        """ok: 0-dim bool / float tensor (loss is finite).  Returns the gradient norm before clipping (device scalar)."""
        flat = self.grads.flat
        flat.copy_(torch.where(ok.to(torch.bool), torch.nan_to_num(flat, nan=0.0, posinf=0.0, neginf=0.0), torch.zeros_like(flat)))
        if reduce:
            self.grads.all_reduce_mean()
        norm = self.grads.clip_(self.cfg.gradient_clip_val) if self.cfg.gradient_clip_val is not None else self.grads.norm()
        self.lr.fill_(self.cfg.lr).mul_(ok.to(self.lr.dtype))
        self.opt.step()
        return norm


class TrainStep:
    """generator + discriminator optimisation of one batch.

    `forward_fn(batch) -> (render_pred, combined_pred)`: the model part (RenderPipeline); `combined_pred.image` is the decoded image
    (b, v, 3, h, w), `render_pred.image` the rendered colour.  `render_losses` / `combined_losses`: LossGroups as configured by
    config/experiment/*.yaml (`loss.target.render.image`, `loss.target.combined`)."""

    def __init__(self, forward_fn: Callable[[dict], tuple], generator_params: Iterable[nn.Parameter], discriminator: Optional[nn.Module],
                 render_losses: LossGroup, combined_losses: LossGroup, last_layer_weights: Optional[Tensor],
                 g_cfg: OptimizerCfg = OptimizerCfg(), d_cfg: OptimizerCfg = OptimizerCfg()):
        self.forward_fn = forward_fn
        self.discriminator = discriminator
        self.render_losses, self.combined_losses = render_losses, combined_losses
        self.last_layer_weights = last_layer_weights
        self.g_opt = GuardedAdam(generator_params, g_cfg)
        self.d_opt = GuardedAdam(discriminator.parameters(), d_cfg) if discriminator is not None else None

    def _logits(self, image: Tensor) -> Tensor:
        b, v = image.shape[:2]
        return self.discriminator(image.flatten(0, 1)).unflatten(0, (b, v))

    def generator_pass(self, batch: dict, global_step: int) -> Dict[str, Tensor]:
        render_pred, combined_pred = self.forward_fn(batch)
        target = batch["target"]
        gt = GroundTruth(image=target["image"], near=target["near"], far=target["far"])
        if self.discriminator is not None and self.combined_losses.is_generator_loss_active(global_step):
            for p in self.discriminator.parameters():          # toggle_optimizer(g_opt): D is applied, not updated (:354, :412-419)
                p.requires_grad_(False)
            combined_pred.logits_fake = self._logits(combined_pred.image)
        a, la = self.render_losses.forward_generator(render_pred, gt, global_step, self.last_layer_weights)
        b, lb = self.combined_losses.forward_generator(combined_pred, gt, global_step, self.last_layer_weights)
        loss = a + b
        self.g_opt.zero()
        if isinstance(loss, Tensor) and loss.requires_grad:
            loss.backward()
            ok = torch.isfinite(loss.detach())
            norm = self.g_opt.step(ok)
        else:                                                  # no active loss at this step
            loss = torch.zeros((), device=self.g_opt.grads.flat.device)
            ok, norm = torch.ones((), dtype=torch.bool, device=loss.device), torch.zeros((), device=loss.device)
        self._last_pred = combined_pred
        if self.discriminator is not None:
            for p in self.discriminator.parameters():
                p.requires_grad_(True)
        out = {"generator/total": loss.detach(), "generator/finite": ok, "generator/grad_norm": norm}
        out.update({f"generator/{k}": v.unweighted.detach() for k, v in {**la, **lb}.items()})
        return out

    def discriminator_pass(self, batch: dict, global_step: int) -> Dict[str, Tensor]:
        if self.discriminator is None or not self.combined_losses.is_discriminator_loss_active(global_step):
            return {}
        pred = Prediction(image=self._last_pred.image.detach())               # NOTE detach (:452)
        pred.logits_fake = self._logits(pred.image)
        pred.logits_real = self._logits(batch["target"]["image"])
        loss, parts = self.combined_losses.forward_discriminator(pred, GroundTruth(image=batch["target"]["image"]), global_step)
        self.d_opt.zero()
        loss.backward()
        ok = torch.isfinite(loss.detach())
        norm = self.d_opt.step(ok)
        out = {"discriminator/total": loss.detach(), "discriminator/finite": ok, "discriminator/grad_norm": norm}
        out.update({f"discriminator/{k}": v.unweighted.detach() for k, v in parts.items()})
        return out

    def __call__(self, batch: dict, global_step: int) -> Dict[str, Tensor]:
        out = self.generator_pass(batch, global_step)
        out.update(self.discriminator_pass(batch, global_step))
        return out
