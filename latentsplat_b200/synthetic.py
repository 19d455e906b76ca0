"""Seeded synthetic inputs for tests and bench.py.

Cameras follow the reference's conventions: extrinsics are camera-to-world, OpenCV axes
(+z forward), intrinsics are normalised by image size
(/root/reference/src/dataset/dataset_re10k.py:139-150, src/geometry/projection.py:233-247).
Everything is generated on the CPU with a torch.Generator so that every rank / box sees the
same numbers; callers move the tensors to the device.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Optional

import torch
from torch import Tensor


def intrinsics(f: float = 0.86) -> Tensor:
    return torch.tensor([[f, 0.0, 0.5], [0.0, f, 0.5], [0.0, 0.0, 1.0]], dtype=torch.float32)


def pose(tx: float = 0.0, yaw_deg: float = 0.0, ty: float = 0.0, tz: float = 0.0) -> Tensor:
    """Camera-to-world: translation (tx,ty,tz) and a yaw about the y axis."""
    a = math.radians(yaw_deg)
    m = torch.eye(4, dtype=torch.float32)
    m[0, 0], m[0, 2], m[2, 0], m[2, 2] = math.cos(a), math.sin(a), -math.sin(a), math.cos(a)
    m[0, 3], m[1, 3], m[2, 3] = tx, ty, tz
    return m


def target_poses(n: int) -> Tensor:
    """Target cameras between/around two context cameras at x=0 and x=1."""
    ts = {1: [0.5], 2: [0.25, 0.75], 3: [0.25, 0.5, 0.75], 4: [-0.25, 0.25, 0.75, 1.25]}.get(
        n, [i / max(n - 1, 1) for i in range(n)])
    return torch.stack([pose(tx=t, yaw_deg=5.0 * (t - 0.5)) for t in ts])


def quaternion_to_matrix(q: Tensor) -> Tensor:
    """Unit quaternion (x, y, z, w) -> rotation matrix."""
    x, y, z, w = q.unbind(-1)
    return torch.stack([1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w),
                        2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w),
                        2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)], dim=-1).reshape(*q.shape[:-1], 3, 3)


@dataclass
class GaussianCloud:
    means: Tensor        # (G,3) world
    covariances: Tensor  # (G,3,3)
    opacities: Tensor    # (G,)


def random_gaussians(G: int, *, seed: int, extrinsics: Optional[Tensor] = None, f: float = 0.86, width: int = 256,
                     near: float = 1.0, far: float = 100.0, s_px: tuple[float, float] = (0.5, 3.0),
                     opacity: tuple[float, float] = (0.02, 0.35)) -> GaussianCloud:
    """Gaussians scattered through the frustum of `extrinsics` (default identity):
    uv ~ U[0,1)^2, depth = 1/U(1/far, 1/near), footprint s_px pixels with U(0.5,2) anisotropy,
    random orientation, opacity ~ U(opacity)."""
    gen = torch.Generator().manual_seed(seed)
    U = lambda *s: torch.rand(*s, generator=gen)
    uv = U(G, 2)
    depth = 1.0 / (1.0 / far + U(G) * (1.0 / near - 1.0 / far))
    cam = torch.stack(((uv[:, 0] - 0.5) / f * depth, (uv[:, 1] - 0.5) / f * depth, depth), dim=-1)
    c2w = extrinsics if extrinsics is not None else torch.eye(4)
    means = cam @ c2w[:3, :3].T + c2w[:3, 3]
    sigma = (s_px[0] + U(G) * (s_px[1] - s_px[0])) * depth / (f * width)
    scales = sigma[:, None] * (0.5 + 1.5 * U(G, 3))
    q = torch.randn(G, 4, generator=gen)
    R = quaternion_to_matrix(q / q.norm(dim=-1, keepdim=True))
    M = R * scales[:, None, :]
    cov = M @ M.transpose(1, 2)
    opac = opacity[0] + U(G) * (opacity[1] - opacity[0])
    return GaussianCloud(means.float(), cov.float(), opac.float())


def sh_mask(degree: int) -> Tensor:
    """DC 1, band l scaled 0.1 * 0.25**l (/root/reference/src/model/encoder/common/gaussian_adapter.py:47-61)."""
    m = torch.ones((degree + 1) ** 2)
    for l in range(1, degree + 1):
        m[l * l:(l + 1) ** 2] = 0.1 * 0.25 ** l
    return m


def random_sh(G: int, channels: int, degree: int, *, seed: int) -> Tensor:
    """(G, channels, (degree+1)^2) coefficients ~ N(0,1) * sh_mask."""
    gen = torch.Generator().manual_seed(seed)
    return torch.randn(G, channels, (degree + 1) ** 2, generator=gen) * sh_mask(degree)
