"""Splatting decoder: Gaussians + target cameras -> colour / feature posterior / mask / depth.

Same constructor, forward signature and outputs as
/root/reference/src/model/decoder/decoder_splatting_cuda.py:20-119; the per-view `repeat`
copies (:71-86) and the per-view rasterizer loop are replaced by one batched call with
`views_per_scene = v`.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Literal, Optional

import torch
from torch import Tensor

from ..diagonal_gaussian_distribution import DiagonalGaussianDistribution
from ..types import Gaussians
from ...rasterizer import RasterDebug
from .cuda_splatting import DepthRenderingMode, RenderOutput, render_cuda, render_depth_cuda
from .decoder import Decoder, DecoderOutput


@dataclass
class DecoderSplattingCUDACfg:
    name: Literal["splatting_cuda"]


class DecoderSplattingCUDA(Decoder[DecoderSplattingCUDACfg]):
    background_color: Tensor

    def __init__(self, cfg: DecoderSplattingCUDACfg, background_color: list[float] = [0.0, 0.0, 0.0],
                 variational: bool = False) -> None:
        super().__init__(cfg)
        self.register_buffer("background_color", torch.tensor(background_color, dtype=torch.float32),
                             persistent=False)
        self.variational = variational
        # Key-list sizing of the rasterizer: None = exact (one 4-byte host sync per forward, always right);
        # an int = sync-free / CUDA-graph capturable with that many (tile, Gaussian) slots.  `last_raster`
        # keeps the device-side counters of the most recent forward so that the caller can verify
        # `last_raster.stats[2] == 0` (no overflow) whenever convenient.
        self.raster_capacity: Optional[int] = None
        self.last_raster = RasterDebug()

    def calibrate_raster_capacity(self, slack: float = 1.5) -> int:
        """After at least one exact forward: capacity = slack x the size that forward needed."""
        if self.last_raster.num_rendered is None:
            raise RuntimeError("run one forward with raster_capacity=None before calibrating")
        self.raster_capacity = int(self.last_raster.num_rendered * slack) + 1024
        return self.raster_capacity

    def render_to_decoder_output(self, render_output: RenderOutput, b: int, v: int) -> DecoderOutput:
        split = lambda t: t.reshape(b, v, *t.shape[1:])
        posterior = None
        if render_output.feature is not None:
            features = split(render_output.feature)
            if self.variational:
                mean, logvar = features.chunk(2, dim=2)
            else:
                # background feature = 0 = mean = logvar of the normal distribution (:45-47)
                mean = features
                logvar = (1 - split(render_output.mask.detach())[:, :, None]).log().expand_as(features)
            posterior = DiagonalGaussianDistribution(mean, logvar)
        return DecoderOutput(color=None if render_output.color is None else split(render_output.color),
                             feature_posterior=posterior, mask=split(render_output.mask),
                             depth=split(render_output.depth))

    def forward(self, gaussians: Gaussians, extrinsics: Tensor, intrinsics: Tensor, near: Tensor, far: Tensor,
                image_shape: tuple[int, int], depth_mode: Optional[DepthRenderingMode] = None,
                return_colors: bool = True, return_features: bool = True) -> DecoderOutput:
        b, v = extrinsics.shape[:2]
        color_sh = gaussians.color_harmonics if return_colors else None
        feature_sh = gaussians.feature_harmonics if return_features else None
        rendered = render_cuda(
            extrinsics.reshape(b * v, 4, 4), intrinsics.reshape(b * v, 3, 3), near.reshape(b * v),
            far.reshape(b * v), image_shape, self.background_color.expand(b * v, 3), gaussians.means,
            gaussians.covariances, gaussians.opacities, color_sh, feature_sh, views_per_scene=v,
            debug=self.last_raster, capacity=self.raster_capacity)
        out = self.render_to_decoder_output(rendered, b, v)
        if depth_mode is not None and depth_mode != "depth":
            out.depth = self.render_depth(gaussians, extrinsics, intrinsics, near, far, image_shape, depth_mode)
        return out

    def render_depth(self, gaussians: Gaussians, extrinsics: Tensor, intrinsics: Tensor, near: Tensor, far: Tensor,
                     image_shape: tuple[int, int], mode: DepthRenderingMode = "depth") -> Tensor:
        b, v = extrinsics.shape[:2]
        result = render_depth_cuda(
            extrinsics.reshape(b * v, 4, 4), intrinsics.reshape(b * v, 3, 3), near.reshape(b * v),
            far.reshape(b * v), image_shape, gaussians.means, gaussians.covariances, gaussians.opacities,
            mode=mode, views_per_scene=v)
        return result.reshape(b, v, *result.shape[1:])

    @property
    def last_layer_weights(self) -> None:
        return None
