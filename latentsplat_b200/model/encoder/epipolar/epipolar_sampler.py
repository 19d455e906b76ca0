"""Samples features of every OTHER context view along each ray's epipolar segment
(/root/reference/src/model/encoder/epipolar/epipolar_sampler.py:18-167)."""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional

import torch
import torch.nn.functional as F
from torch import Tensor, nn

from ....geometry.epipolar_lines import project_rays
from ....geometry.projection import get_world_rays, sample_image_grid
from ....misc.heterogeneous_pairings import generate_heterogeneous_index, generate_heterogeneous_index_transpose


@dataclass
class EpipolarSampling:
    features: Optional[Tensor]   # (batch, view, other_view, ray, sample, channel); None when the caller gathers (fused path)
    valid: Tensor           # (batch, view, other_view, ray) bool
    xy_ray: Tensor          # (batch, view, ray, 2)
    xy_sample: Tensor       # (batch, view, other_view, ray, sample, 2)
    xy_sample_near: Tensor
    xy_sample_far: Tensor
    origins: Tensor         # (batch, view, ray, 3)
    directions: Tensor      # (batch, view, ray, 3)


class EpipolarSampler(nn.Module):
    def __init__(self, num_views: int, num_samples: int) -> None:
        super().__init__()
        self.num_samples = num_samples
        _, index_v = generate_heterogeneous_index(num_views)
        t_v, t_ov = generate_heterogeneous_index_transpose(num_views)
        self.register_buffer("index_v", index_v, persistent=False)
        self.register_buffer("transpose_v", t_v, persistent=False)
        self.register_buffer("transpose_ov", t_ov, persistent=False)

    def forward(self, images: Tensor, extrinsics: Tensor, intrinsics: Tensor, near: Tensor, far: Tensor,
                gather: bool = True) -> EpipolarSampling:
        """`gather=False` returns the sampling geometry only (`features=None`): the caller fuses the bilinear gather, the
        overlap mask and the depth encoding into one kernel (latentsplat_b200.epipolar_gather) using `image_index`."""
        b, v, c, h, w = images.shape
        device = images.device
        xy_ray, origins, directions = self.generate_image_rays(images, extrinsics, intrinsics)
        projection = project_rays(origins[:, :, None], directions[:, :, None],
                                  self.collect(extrinsics)[:, :, :, None], self.collect(intrinsics)[:, :, :, None],
                                  near[:, :, None, None], far[:, :, None, None])
        s = self.num_samples
        sample_depth = ((torch.arange(s, device=device) + 0.5) / s)[:, None]
        overlaps = projection["overlaps_image"]
        xy_min = (projection["xy_min"].nan_to_num(posinf=0, neginf=0) * overlaps[..., None])[..., None, :]
        xy_max = (projection["xy_max"].nan_to_num(posinf=0, neginf=0) * overlaps[..., None])[..., None, :]
        xy_sample = xy_min + sample_depth * (xy_max - xy_min)
        half_span = 0.5 / s
        if not gather:
            return EpipolarSampling(features=None, valid=overlaps, xy_ray=xy_ray, xy_sample=xy_sample,
                                    xy_sample_near=xy_min + (sample_depth - half_span) * (xy_max - xy_min),
                                    xy_sample_far=xy_min + (sample_depth + half_span) * (xy_max - xy_min),
                                    origins=origins, directions=directions)

        # Before the "transpose", dim 1 is the view the ray is cast from; sampling needs the view drawn from.
        samples = self.transpose(xy_sample)
        ov = v - 1
        samples = F.grid_sample(images.flatten(0, 1), (2 * samples - 1).reshape(b * v, ov * xy_ray.shape[2] * s, 1, 2),
                                mode="bilinear", padding_mode="zeros", align_corners=False)
        samples = samples[..., 0].transpose(1, 2).reshape(b, v, ov, -1, s, c)     # (b v) c (ov r s) () -> b v ov r s c
        samples = self.transpose(samples)
        samples = samples * overlaps[..., None, None]                              # zero out invalid samples

        return EpipolarSampling(features=samples, valid=overlaps, xy_ray=xy_ray, xy_sample=xy_sample,
                                xy_sample_near=xy_min + (sample_depth - half_span) * (xy_max - xy_min),
                                xy_sample_far=xy_min + (sample_depth + half_span) * (xy_max - xy_min),
                                origins=origins, directions=directions)

    def generate_image_rays(self, images: Tensor, extrinsics: Tensor, intrinsics: Tensor):
        """One ray per pixel of the (downscaled) feature map."""
        b, v, _, h, w = images.shape
        xy, _ = sample_image_grid((h, w), device=images.device)
        xy = xy.reshape(h * w, 2)
        origins, directions = get_world_rays(xy, extrinsics[:, :, None], intrinsics[:, :, None])
        return xy[None, None].expand(b, v, -1, -1), origins, directions

    def image_index(self, b: int, rays: int) -> Tensor:
        """(b * v * (v-1) * rays) int32: flattened (scene * views + other view) feature map each epipolar line lies in --
        what the two `transpose` shuffles around grid_sample express (:96-110)."""
        key = (b, rays)
        if getattr(self, "_image_index_key", None) != key or self._image_index.device != self.index_v.device:
            v = self.index_v.shape[0]
            idx = torch.arange(b, device=self.index_v.device)[:, None, None] * v + self.index_v[None]        # (b, v, ov)
            self._image_index = idx[..., None].expand(b, v, v - 1, rays).reshape(-1).to(torch.int32).contiguous()
            self._image_index_key = key
        return self._image_index

    def transpose(self, x: Tensor) -> Tensor:
        b, v, ov = x.shape[:3]
        t_b = torch.arange(b, device=x.device)[:, None, None].expand(b, v, ov)
        return x[t_b, self.transpose_v[None].expand(b, v, ov), self.transpose_ov[None].expand(b, v, ov)]

    def collect(self, target: Tensor) -> Tensor:
        """(batch, view, …) -> (batch, view, view-1, …): for each view, the other views."""
        b, v = target.shape[:2]
        index_b = torch.arange(b, device=target.device)[:, None, None].expand(b, v, v - 1)
        return target[index_b, self.index_v[None].expand(b, v, v - 1)]
