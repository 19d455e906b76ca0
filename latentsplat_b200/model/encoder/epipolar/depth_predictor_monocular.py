"""Per-ray depth distribution -> sampled depths and densities
(/root/reference/src/model/encoder/epipolar/depth_predictor_monocular.py:10-81)."""
from __future__ import annotations

import torch
from torch import Tensor, nn

from .conversions import relative_disparity_to_depth
from .distribution_sampler import DistributionSampler
from latentsplat_b200.gemm import Linear  # nn.Linear with tcgen05 TF32 GEMMs on CUDA


class DepthPredictorMonocular(nn.Module):
    def __init__(self, d_in: int, num_samples: int, num_surfaces: int, use_transmittance: bool) -> None:
        super().__init__()
        self.projection = nn.Sequential(nn.ReLU(), Linear(d_in, 2 * num_samples * num_surfaces))
        self.sampler = DistributionSampler()
        self.num_samples = num_samples
        self.num_surfaces = num_surfaces
        self.use_transmittance = use_transmittance
        self.to_pdf = nn.Softmax(dim=-1)      # modules so that hooks can latch on (:33-35)
        self.to_offset = nn.Sigmoid()

    def forward(self, features: Tensor, near: Tensor, far: Tensor, deterministic: bool, gaussians_per_pixel: int):
        """features (b, v, ray, c) -> depth, density, each (b, v, ray, surface, sample)."""
        s = self.num_samples
        x = self.projection(features)
        # "... (dpt srf c) -> c ... srf dpt"
        x = x.unflatten(-1, (s, self.num_surfaces, 2)).movedim(-1, 0).transpose(-1, -2)
        pdf, offset = self.to_pdf(x[0]), self.to_offset(x[1])
        index, pdf_i = self.sampler.sample(pdf, deterministic, gaussians_per_pixel)
        offset = self.sampler.gather(index, offset)
        relative_disparity = (index + offset) / s
        depth = relative_disparity_to_depth(relative_disparity, near[..., None, None, None], far[..., None, None, None])
        if self.use_transmittance:
            partial = pdf.cumsum(dim=-1)
            partial = torch.cat((torch.zeros_like(partial[..., :1]), partial[..., :-1]), dim=-1)
            opacity = self.sampler.gather(index, pdf / (1 - partial + 1e-10))
        else:
            opacity = pdf_i
        return depth, opacity
