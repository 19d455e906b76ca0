"""Raw per-pixel network outputs -> world-space Gaussians
(/root/reference/src/model/encoder/common/gaussian_adapter.py:13-139)."""
from __future__ import annotations

from dataclasses import dataclass

import torch
from torch import Tensor, nn

from ....geometry.projection import get_world_rays
from ....misc.cache import device_constant
from ....misc.sh_utils import sh_rotation_matrices
from .gaussians import build_covariance, outer_sym, quaternion_to_matrix  # noqa: F401
from latentsplat_b200.geometry.inverse import inv2x2  # closed-form camera inverses (no cuSOLVER)


@dataclass
class Gaussians:
    means: Tensor             # (*batch, 3)
    covariances: Tensor       # (*batch, 3, 3)
    scales: Tensor            # (*batch, 3)
    rotations: Tensor         # (*batch, 4)
    color_harmonics: Tensor   # (*batch, 3, d_color_sh)
    feature_harmonics: Tensor # (*batch, channels, d_feature_sh)
    opacities: Tensor         # (*batch)


@dataclass
class GaussianAdapterCfg:
    gaussian_scale_min: float
    gaussian_scale_max: float
    color_sh_degree: int
    feature_sh_degree: int


class GaussianAdapter(nn.Module):
    def __init__(self, cfg: GaussianAdapterCfg, n_feature_channels: int):
        super().__init__()
        self.cfg = cfg
        self.n_feature_channels = n_feature_channels
        # DC coefficient 1, band l scaled by 0.1 * 0.25**l: large DC, small view-dependent part at init (:44-61)
        for name, degree in (("color_sh_mask", cfg.color_sh_degree), ("feature_sh_mask", cfg.feature_sh_degree)):
            mask = torch.ones(((degree + 1) ** 2,), dtype=torch.float32)
            for l in range(1, degree + 1):
                mask[l * l:(l + 1) ** 2] = 0.1 * 0.25 ** l
            self.register_buffer(name, mask, persistent=False)

    def harmonics_transform(self, c2w_rotations: Tensor) -> Tensor:
        """(…, 3, 3) camera-to-world rotations -> (…, d_in, d_in) matrix T with  T @ raw  ==  raw with its colour and
        feature SH blocks masked (:44-61) and rotated into world space (:104-105), everything else untouched.
        Masking and rotation are linear in the raw network output, so the encoder folds T into the weights of the
        `to_gaussians` Linear (one effective weight per context view) and calls `forward(..., harmonics_ready=True)`:
        the two mask multiplies, the per-degree batched rotations over millions of coefficient vectors, their `cat`s
        and all of their backward passes disappear into a GEMM that runs anyway."""
        lead = c2w_rotations.shape[:-2]
        t = torch.zeros((*lead, self.d_in, self.d_in), dtype=c2w_rotations.dtype, device=c2w_rotations.device)
        torch.diagonal(t[..., :7, :7], dim1=-2, dim2=-1).fill_(1.0)      # scales, rotation: untouched (no H2D: capturable)
        offset = 7
        for degree, mask, channels, width in ((self.cfg.color_sh_degree, self.color_sh_mask, 3, self.d_color_sh),
                                              (self.cfg.feature_sh_degree, self.feature_sh_mask, self.n_feature_channels,
                                               self.d_feature_sh)):
            mats = sh_rotation_matrices(c2w_rotations, degree)
            block = torch.zeros((*lead, width, width), dtype=t.dtype, device=t.device)
            for l, m in enumerate(mats):
                block[..., l * l:(l + 1) ** 2, l * l:(l + 1) ** 2] = m
            block = block * mask                                   # D diag(mask): scale column j by mask[j]
            for ch in range(channels):
                t[..., offset:offset + width, offset:offset + width] = block
                offset += width
        return t

    def forward(self, extrinsics: Tensor, intrinsics: Tensor, coordinates: Tensor, depths: Tensor, opacities: Tensor,
                raw_gaussians: Tensor, image_shape: tuple[int, int], eps: float = 1e-8,
                harmonics_ready: bool = False) -> Gaussians:
        device = extrinsics.device
        scales, rotations, color_sh, feature_sh = raw_gaussians.split(
            (3, 4, 3 * self.d_color_sh, self.n_feature_channels * self.d_feature_sh), dim=-1)

        scale_min, scale_max = self.cfg.gaussian_scale_min, self.cfg.gaussian_scale_max
        scales = scale_min + (scale_max - scale_min) * scales.sigmoid()
        h, w = image_shape
        pixel_size = device_constant((1 / w, 1 / h), device)
        scales = scales * depths[..., None] * self.get_scale_multiplier(intrinsics, pixel_size)[..., None]

        rotations = rotations / (rotations.norm(dim=-1, keepdim=True) + eps)

        color_sh = color_sh.unflatten(-1, (3, self.d_color_sh))
        feature_sh = feature_sh.unflatten(-1, (self.n_feature_channels, self.d_feature_sh))
        if not harmonics_ready:
            color_sh = color_sh * self.color_sh_mask
            feature_sh = feature_sh * self.feature_sh_mask

        # world-space covariance C (R S S^T R^T) C^T = (C R S)(C R S)^T, without batched 3x3 GEMMs
        c2w_rotations = extrinsics[..., :3, :3]
        local = quaternion_to_matrix(rotations) * scales[..., None, :]
        covariances = outer_sym((c2w_rotations[..., :, :, None] * local[..., None, :, :]).sum(dim=-2))

        origins, directions = get_world_rays(coordinates, extrinsics, intrinsics)
        means = origins + directions * depths[..., None]

        if not harmonics_ready:
            # rotate once per ray, THEN broadcast over the samples of the ray (the reference broadcasts first, :92-93,
            # and rotates spp x as many coefficient vectors)
            color_sh = self._rotate(color_sh, c2w_rotations, self.cfg.color_sh_degree)
            feature_sh = self._rotate(feature_sh, c2w_rotations, self.cfg.feature_sh_degree)
        return Gaussians(means=means, covariances=covariances,
                         color_harmonics=color_sh.broadcast_to((*opacities.shape, 3, self.d_color_sh)),
                         feature_harmonics=feature_sh.broadcast_to((*opacities.shape, self.n_feature_channels,
                                                                    self.d_feature_sh)),
                         opacities=opacities,
                         scales=scales,                                             # camera space (ply export only)
                         rotations=rotations.broadcast_to((*scales.shape[:-1], 4)))

    @staticmethod
    def _rotate(sh: Tensor, c2w_rotations: Tensor, degree: int) -> Tensor:
        """rotate_sh(sh, c2w[..., None, :, :]) (:104-105): the rotation matrices are built once per distinct
        camera (a few per batch) and applied per degree block."""
        mats = sh_rotation_matrices(c2w_rotations[..., None, :, :], degree)
        parts = [torch.einsum("...ij,...j->...i", mats[l], sh[..., l * l:(l + 1) ** 2]) for l in range(degree + 1)]
        return torch.cat(parts, dim=-1)

    def get_scale_multiplier(self, intrinsics: Tensor, pixel_size: Tensor, multiplier: float = 0.1) -> Tensor:
        inv = inv2x2(intrinsics[..., :2, :2])
        return (multiplier * torch.einsum("...ij,j->...i", inv, pixel_size)).sum(dim=-1)

    @property
    def d_color_sh(self) -> int:
        return (self.cfg.color_sh_degree + 1) ** 2

    @property
    def d_feature_sh(self) -> int:
        return (self.cfg.feature_sh_degree + 1) ** 2

    @property
    def d_in(self) -> int:
        return 7 + 3 * self.d_color_sh + self.n_feature_channels * self.d_feature_sh
