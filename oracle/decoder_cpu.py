"""CPU restatement of the splatting decoder (TEST INFRASTRUCTURE / reference arm of bench.py only).

Follows /root/reference/src/model/decoder/cuda_splatting.py:56-167 and decoder_splatting_cuda.py:37-91 literally:
per-view repeat of the Gaussians, 1/near rescale in torch, feature SH evaluated in torch (0.5 + eval_sh), one
rasterizer object per view (here the CPU oracle behind the reference's call signature), stacked outputs.
Never imported by latentsplat_b200/.
"""
from __future__ import annotations

from math import isqrt

import torch
from torch import nn

from oracle.raster_stub import OracleGaussianRasterizer, OracleSettings


def _eval_sh(deg, sh, dirs):
    """src/misc/sh_utils.py:42-97 (deg <= 2 is all the feature path needs; deg 4 goes through the rasterizer)."""
    C0, C1 = 0.28209479177387814, 0.4886025119029199
    C2 = [1.0925484305920792, -1.0925484305920792, 0.31539156525252005, -1.0925484305920792, 0.5462742152960396]
    result = C0 * sh[..., 0]
    if deg > 0:
        x, y, z = dirs[..., 0:1], dirs[..., 1:2], dirs[..., 2:3]
        result = result - C1 * x * sh[..., 1] + C1 * y * sh[..., 2] - C1 * z * sh[..., 3]
        if deg > 1:
            xx, yy, zz, xy, yz, xz = x * x, y * y, z * z, x * y, y * z, x * z
            result = (result + C2[0] * xz * sh[..., 4] + C2[1] * xy * sh[..., 5] + C2[2] * (2.0 * yy - zz - xx) * sh[..., 6]
                      + C2[3] * yz * sh[..., 7] + C2[4] * (zz - xx) * sh[..., 8])
    assert deg <= 2
    return result


def render_cpu(extrinsics, intrinsics, near, far, image_shape, background_color, means, covariances, opacities,
               color_sh=None, feature_sh=None):
    from latentsplat_b200.model.decoder.cuda_splatting import get_fov, get_projection_matrix
    scale = 1 / near
    extrinsics = extrinsics.clone()
    extrinsics[..., :3, 3] = extrinsics[..., :3, 3] * scale[:, None]
    covariances = covariances * (scale[:, None, None, None] ** 2)
    means = means * scale[:, None, None]
    near, far = near * scale, far * scale
    shs = features = None
    deg = 0
    if color_sh is not None:
        deg = isqrt(color_sh.shape[-1]) - 1
        shs = color_sh.transpose(-1, -2).contiguous()
    if feature_sh is not None:
        d = means - extrinsics[:, None, :3, 3]
        features = 0.5 + _eval_sh(isqrt(feature_sh.shape[-1]) - 1, feature_sh, d / d.norm(dim=-1, keepdim=True))
    b = extrinsics.shape[0]
    h, w = image_shape
    fov_x, fov_y = get_fov(intrinsics).unbind(dim=-1)
    tan_x, tan_y = (0.5 * fov_x).tan(), (0.5 * fov_y).tan()
    proj = get_projection_matrix(near, far, fov_x, fov_y).transpose(1, 2)
    view = extrinsics.inverse().transpose(1, 2)
    full = view @ proj
    row, col = torch.triu_indices(3, 3)
    outs = []
    for i in range(b):
        settings = OracleSettings(h, w, tan_x[i].item(), tan_y[i].item(), background_color[i], 1.0, view[i], full[i], deg,
                                  extrinsics[i, :3, 3], False, False)
        outs.append(OracleGaussianRasterizer(settings)(
            means3D=means[i], means2D=torch.zeros_like(means[i], requires_grad=True),
            shs=None if shs is None else shs[i], colors_precomp=None, features=None if features is None else features[i],
            opacities=opacities[i, ..., None], cov3D_precomp=covariances[i, :, row, col]))
    stack = lambda k: None if outs[0][k] is None else torch.stack([o[k] for o in outs])
    return stack(0), stack(1), stack(2)[:, 0], stack(3)[:, 0]


class DecoderSplattingCPU(nn.Module):
    """Drop-in for DecoderSplattingCUDA on CPU tensors (reference data flow, oracle rasterizer)."""

    def __init__(self, background_color=(0.0, 0.0, 0.0), variational: bool = False, n_threads: int = 0):
        super().__init__()
        self.register_buffer("background_color", torch.tensor(background_color, dtype=torch.float32), persistent=False)
        self.variational = variational
        OracleGaussianRasterizer.n_threads = n_threads

    def forward(self, gaussians, extrinsics, intrinsics, near, far, image_shape, depth_mode=None, return_colors=True,
                return_features=True):
        from latentsplat_b200.model.decoder.decoder import DecoderOutput
        from latentsplat_b200.model.diagonal_gaussian_distribution import DiagonalGaussianDistribution
        b, v = extrinsics.shape[:2]
        rep = lambda t: None if t is None else t.repeat_interleave(v, dim=0)
        color, feature, mask, depth = render_cpu(
            extrinsics.flatten(0, 1), intrinsics.flatten(0, 1), near.flatten(), far.flatten(), image_shape,
            self.background_color.expand(b * v, 3), rep(gaussians.means), rep(gaussians.covariances),
            rep(gaussians.opacities), rep(gaussians.color_harmonics) if return_colors else None,
            rep(gaussians.feature_harmonics) if return_features else None)
        split = lambda t: None if t is None else t.unflatten(0, (b, v))
        posterior = None
        if feature is not None:
            f = split(feature)
            if self.variational:
                mean, logvar = f.chunk(2, dim=2)
            else:
                mean, logvar = f, (1 - split(mask.detach())[:, :, None]).log().expand_as(f)
            posterior = DiagonalGaussianDistribution(mean, logvar)
        return DecoderOutput(split(color), posterior, split(mask), split(depth))
