"""Closed-form inverses of the small camera matrices on the hot path.

The reference calls `.inverse()` on (b, v, 4, 4) extrinsics and (b, v, 3, 3) intrinsics
(/root/reference/src/model/decoder/cuda_splatting.py:40, 79; src/geometry/projection.py:39, 59;
src/geometry/epipolar_lines.py:70; src/model/encoder/common/gaussian_adapter.py:122): on CUDA that is a batched LU through
cuSOLVER / MAGMA-style kernels (getrf, laswp, two trsm) per call, with an info tensor on the side.  These matrices are 2x2, 3x3 or
affine 4x4 with the last row (0, 0, 0, 1), so the adjugate formula is exact, sync-free, differentiable and a handful of elementwise
kernels that fuse into the surrounding CUDA graph.
"""
from __future__ import annotations

import torch
from torch import Tensor


def inv2x2(m: Tensor) -> Tensor:
    a, b, c, d = m[..., 0, 0], m[..., 0, 1], m[..., 1, 0], m[..., 1, 1]
    det = a * d - b * c
    return torch.stack((torch.stack((d, -b), -1), torch.stack((-c, a), -1)), -2) / det[..., None, None]


def inv3x3(m: Tensor) -> Tensor:
    """inverse = [r1 x r2, r2 x r0, r0 x r1] as COLUMNS, divided by det = r0 . (r1 x r2)."""
    r0, r1, r2 = m[..., 0, :], m[..., 1, :], m[..., 2, :]
    c0 = torch.linalg.cross(r1, r2, dim=-1)
    c1 = torch.linalg.cross(r2, r0, dim=-1)
    c2 = torch.linalg.cross(r0, r1, dim=-1)
    det = (r0 * c0).sum(-1)
    return torch.stack((c0, c1, c2), dim=-1) / det[..., None, None]


def inv_affine4x4(m: Tensor) -> Tensor:
    """[[A, t], [0, 1]]^-1 = [[A^-1, -A^-1 t], [0, 1]] (camera poses: rigid, possibly with a scale; the last row is (0,0,0,1))."""
    a_inv = inv3x3(m[..., :3, :3])
    t = -(a_inv @ m[..., :3, 3:4])
    top = torch.cat((a_inv, t), dim=-1)
    bottom = torch.zeros_like(m[..., 3:4, :])
    bottom[..., 0, 3] = 1.0
    return torch.cat((top, bottom), dim=-2)
