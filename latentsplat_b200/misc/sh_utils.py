"""Real spherical harmonics of degree <= 4 in the reference's basis, and their rotation.

`eval_sh` restates /root/reference/src/misc/sh_utils.py:42-97 (pinned by tests/golden/sh_eval.npz).

`rotate_sh` replaces the reference's e3nn-based version (sh_utils.py:100-120: `matrix_to_angles` +
`wigner_D`, e3nn==0.5.1, not installed here -- PARITY UNPINNED against e3nn).  It uses the defining
property instead of Euler angles: the coefficients c' of the function rotated by R satisfy
    sum_k c'_k Y_k(d) = sum_k c_k Y_k(R^T d)        for every direction d,
and, since each degree's basis functions are linearly independent, the matrix with that property is unique:
    D_l(R) = pinv(Y_l(S)) @ Y_l(S R)                 (S = fixed sample directions, rows; Y_l(S R)_i = Y_l(R^T s_i))
which is exactly the Wigner-D matrix of the basis `eval_sh` uses.  No trigonometry, no matrix exponentials,
batched over rotations; tests check the defining identity, orthogonality and the group law.
"""
from __future__ import annotations

from math import isqrt

import torch
from torch import Tensor

C0 = 0.28209479177387814
C1 = 0.4886025119029199
C2 = (1.0925484305920792, -1.0925484305920792, 0.31539156525252005, -1.0925484305920792, 0.5462742152960396)
C3 = (-0.5900435899266435, 2.890611442640554, -0.4570457994644658, 0.3731763325901154, -0.4570457994644658,
      1.445305721320277, -0.5900435899266435)
C4 = (2.5033429417967046, -1.7701307697799304, 0.9461746957575601, -0.6690465435572892, 0.10578554691520431,
      -0.6690465435572892, 0.47308734787878004, -1.7701307697799304, 0.6258357354491761)


def sh_basis(deg: int, dirs: Tensor, harmonic: bool = False) -> Tensor:
    """(…,3) unit directions -> (…, (deg+1)^2) basis values; basis[k] multiplies coefficient k.

    `harmonic=False` reproduces the reference's `eval_sh` literally, INCLUDING its coefficient 14, which is
    written `z (zz - xx)` (sh_utils.py:84) where the degree-3 harmonic is `y (zz - xx)` (stock 3DGS
    `z (xx - yy)` under the reference's axis relabelling x<-z, y<-x, z<-y).  As written, function 14 is not
    orthogonal to functions 3, 13, 15 and the degree-3 block is not closed under rotation.  Rendering must
    keep the reference's polynomial (pinned by tests/golden/sh_eval.npz); `harmonic=True` gives the true
    harmonic, which is what a Wigner-D matrix (e3nn) rotates -- used only to derive rotation matrices."""
    assert 0 <= deg <= 4
    x, y, z = dirs[..., 0], dirs[..., 1], dirs[..., 2]
    out = [torch.full_like(x, C0)]
    if deg > 0:
        out += [-C1 * x, C1 * y, -C1 * z]
    if deg > 1:
        xx, yy, zz, xy, yz, xz = x * x, y * y, z * z, x * y, y * z, x * z
        out += [C2[0] * xz, C2[1] * xy, C2[2] * (2.0 * yy - zz - xx), C2[3] * yz, C2[4] * (zz - xx)]
    if deg > 2:
        out += [C3[0] * x * (3 * zz - xx), C3[1] * xz * y, C3[2] * x * (4 * yy - zz - xx),
                C3[3] * y * (2 * yy - 3 * zz - 3 * xx), C3[4] * z * (4 * yy - zz - xx),
                C3[5] * (y if harmonic else z) * (zz - xx),
                C3[6] * z * (zz - 3 * xx)]
    if deg > 3:
        out += [C4[0] * xz * (zz - xx), C4[1] * xy * (3 * zz - xx), C4[2] * xz * (7 * yy - 1),
                C4[3] * xy * (7 * yy - 3), C4[4] * (yy * (35 * yy - 30) + 3), C4[5] * yz * (7 * yy - 3),
                C4[6] * (zz - xx) * (7 * yy - 1), C4[7] * yz * (zz - 3 * xx),
                C4[8] * (zz * (zz - 3 * xx) - xx * (3 * zz - xx))]
    return torch.stack(out, dim=-1)


def eval_sh(deg: int, sh: Tensor, dirs: Tensor) -> Tensor:
    """sh (…, C, >= (deg+1)^2), dirs (…, 3) unit -> (…, C)."""
    n = (deg + 1) ** 2
    assert sh.shape[-1] >= n
    return (sh[..., :n] * sh_basis(deg, dirs)[..., None, :]).sum(dim=-1)


_SAMPLES: dict = {}


def _sample_set(device, dtype):
    """Fixed, well-spread sample directions and the per-degree pseudo-inverses of their basis matrices."""
    key = (str(device), dtype)
    if key not in _SAMPLES:
        n = 48
        i = torch.arange(n, dtype=torch.float64) + 0.5
        phi = torch.acos(1 - 2 * i / n)                     # Fibonacci sphere
        theta = torch.pi * (1 + 5 ** 0.5) * i
        s = torch.stack((torch.cos(theta) * torch.sin(phi), torch.sin(theta) * torch.sin(phi), torch.cos(phi)), -1)
        basis = sh_basis(4, s, harmonic=True)
        pinvs = [torch.linalg.pinv(basis[:, l * l:(l + 1) ** 2]).to(device=device, dtype=dtype) for l in range(5)]
        _SAMPLES[key] = (s.to(device=device, dtype=dtype), pinvs)
    return _SAMPLES[key]


def sh_rotation_matrices(rotations: Tensor, max_degree: int) -> list[Tensor]:
    """rotations (…,3,3) -> [D_0 (…,1,1), D_1 (…,3,3), …] with c'_l = D_l c_l."""
    s, pinvs = _sample_set(rotations.device, rotations.dtype)
    rotated = torch.einsum("sj,...ji->...si", s, rotations)          # rows R^T s_i
    basis = sh_basis(max_degree, rotated, harmonic=True)             # (…, S, n)
    return [pinvs[l] @ basis[..., l * l:(l + 1) ** 2] for l in range(max_degree + 1)]


def rotate_sh(sh_coefficients: Tensor, rotations: Tensor) -> Tensor:
    """sh_coefficients (*#batch, n), rotations (*#batch, 3, 3) -> rotated coefficients (*batch, n)."""
    n = sh_coefficients.shape[-1]
    deg = isqrt(n) - 1
    mats = sh_rotation_matrices(rotations, deg)
    parts = [torch.einsum("...ij,...j->...i", mats[l], sh_coefficients[..., l * l:(l + 1) ** 2])
             for l in range(deg + 1)]
    return torch.cat(parts, dim=-1)
