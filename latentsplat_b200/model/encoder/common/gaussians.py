"""Covariance from scale + quaternion (/root/reference/src/model/encoder/common/gaussians.py:8-44)."""
from __future__ import annotations

import torch
from torch import Tensor


def quaternion_to_matrix(quaternions: Tensor, eps: float = 1e-8) -> Tensor:
    """(…,4) quaternions in (x, y, z, w) order -> (…,3,3)."""
    i, j, k, r = torch.unbind(quaternions, dim=-1)
    two_s = 2 / ((quaternions * quaternions).sum(dim=-1) + eps)
    o = torch.stack((1 - two_s * (j * j + k * k), two_s * (i * j - k * r), two_s * (i * k + j * r),
                     two_s * (i * j + k * r), 1 - two_s * (i * i + k * k), two_s * (j * k - i * r),
                     two_s * (i * k - j * r), two_s * (j * k + i * r), 1 - two_s * (i * i + j * j)), -1)
    return o.unflatten(-1, (3, 3))


def outer_sym(m: Tensor) -> Tensor:
    """m @ m^T for (…,3,3) written as elementwise sums: a batched 3x3 `matmul` over millions of Gaussians lands
    on cuBLAS batched-GEMM tiles of 64x256 (measured 90 ms per call at 1.5 M Gaussians with TF32 enabled)."""
    return (m[..., :, None, :] * m[..., None, :, :]).sum(dim=-1)


def build_covariance(scale: Tensor, rotation_xyzw: Tensor) -> Tensor:
    """R S S^T R^T."""
    rotation = quaternion_to_matrix(rotation_xyzw)
    return outer_sym(rotation * scale[..., None, :])          # (R diag(s)) (R diag(s))^T
