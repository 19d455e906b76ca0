"""The render hot path as one module: encoder -> splatting decoder -> latent sample -> 1/8 rescale -> VAE decode
(-> PatchGAN logits), i.e. the model part of `ModelWrapper.training_step`
(/root/reference/src/model/model_wrapper.py:352-385 and :412-419) without the Lightning harness, loss groups,
logging or optimisers (out of scope).
"""
from __future__ import annotations

from dataclasses import dataclass
from fractions import Fraction
from typing import Optional

import torch
import torch.nn.functional as F
from torch import Tensor, nn

from .model.decoder.decoder import DecoderOutput
from .model.types import VariationalGaussians


@dataclass
class PipelineOutput:
    gaussians: VariationalGaussians
    render: DecoderOutput
    latent_sample: Tensor                 # (b, v, C, H, W) sample of the rendered feature posterior
    z: Tensor                             # (b, v, C, H/s, W/s) after inverting the supersampling
    image: Optional[Tensor]               # (b, v, 3, H, W) VAE-decoded prediction
    logits_fake: Optional[Tensor]         # (b, v, 1, h, w) PatchGAN logits of `image`


def rescale(x: Tensor, scale_factor: Fraction) -> Tensor:
    """Anti-aliased bilinear resize of the trailing two dims (`ModelWrapper.rescale`, model_wrapper.py:266-274:
    torchvision `resize(antialias=True)` == `interpolate(mode="bilinear", antialias=True)` on tensors)."""
    batch_dims, spatial = x.shape[:-2], x.shape[-2:]
    size = tuple(int(Fraction(scale_factor) * s) for s in spatial)
    y = F.interpolate(x.reshape(-1, 1, *spatial), size=size, mode="bilinear", antialias=True, align_corners=False)
    return y.view(*batch_dims, *size)


class RenderPipeline(nn.Module):
    def __init__(self, autoencoder, encoder, decoder, discriminator=None, supersampling_factor: int = 8,
                 variational: str = "gaussians"):
        super().__init__()
        self.autoencoder, self.encoder, self.decoder, self.discriminator = autoencoder, encoder, decoder, discriminator
        self.supersampling_factor = supersampling_factor
        self.variational = variational

    def forward(self, batch: dict, global_step: int = 0, deterministic: bool = False, decode_image: bool = True,
                return_colors: bool = True, discriminate: bool = False) -> PipelineOutput:
        context, target = batch["context"], batch["target"]
        size = tuple(context["image"].shape[-2:]) if "image_shape" not in batch else batch["image_shape"]
        gaussians: VariationalGaussians = self.encoder(context, global_step, features=None, deterministic=deterministic)
        if self.variational not in ("gaussians", "none"):
            g = gaussians.flatten()
        else:                                  # model_wrapper.py:362 (sample) / :630 (mode); no RNG draw when deterministic
            g = gaussians.mode() if deterministic else gaussians.sample()
        out = self.decoder(g, target["extrinsics"], target["intrinsics"], target["near"], target["far"], size,
                           return_colors=return_colors, return_features=True)
        latent_sample = out.feature_posterior.mode() if deterministic else out.feature_posterior.sample()
        z = rescale(latent_sample, Fraction(1, self.supersampling_factor))          # invert supersampling (:376)
        image = logits = None
        if decode_image:
            skip_z = None
            if self.autoencoder.expects_skip:
                skip_z = torch.cat((out.color.detach(), latent_sample), dim=-3) if self.autoencoder.expects_skip_extra \
                    else latent_sample
            image = self.autoencoder.decode(z, skip_z)
            if discriminate and self.discriminator is not None:
                b, v = image.shape[:2]
                logits = self.discriminator(image.flatten(0, 1)).unflatten(0, (b, v))
        return PipelineOutput(gaussians, out, latent_sample, z, image, logits)
