"""Shared test helpers: camera matrices the way the reference builds them, oracle calls, tolerances."""
from __future__ import annotations

import math

import numpy as np
import torch

from latentsplat_b200 import synthetic
from oracle import oracle


def camera(extrinsics: torch.Tensor, f: float, near: float, far: float, fy: float | None = None):
    """view^T, (view^T @ proj^T), campos, tanfov as at cuda_splatting.py:111-118 (numpy float32)."""
    fy = f if fy is None else fy
    tanx, tany = 0.5 / f, 0.5 / fy
    proj = np.zeros((4, 4), np.float32)
    proj[0, 0], proj[1, 1] = 1.0 / tanx, 1.0 / tany
    proj[3, 2] = 1.0
    proj[2, 2] = far / (far - near)
    proj[2, 3] = -(far * near) / (far - near)
    e = extrinsics.double().numpy()
    view_t = np.linalg.inv(e).T.astype(np.float32)
    full_t = (view_t.astype(np.float64) @ proj.T.astype(np.float64)).astype(np.float32)
    return dict(viewmatrix=view_t, projmatrix=full_t, campos=e[:3, 3].astype(np.float32), tanfovx=tanx, tanfovy=tany)


def cov6(cov: torch.Tensor) -> torch.Tensor:
    i, j = torch.triu_indices(3, 3)
    return cov[..., i, j].contiguous()


def raster_case(G=2000, H=64, W=64, seed=1, f=0.86, C=4, color="sh", sh_degree=4, near=1.0, far=100.0,
                extrinsics=None, s_px=(0.5, 3.0), opacity=(0.02, 0.35)):
    """One view of random Gaussians + camera (everything numpy float32)."""
    view = extrinsics if extrinsics is not None else synthetic.pose(0.1, 3.0)
    cloud = synthetic.random_gaussians(G, seed=seed, f=f, width=W, near=near, far=far, s_px=s_px, opacity=opacity)
    cam = camera(view, f, near, far)
    d = dict(means3D=cloud.means.numpy(), cov3D=cov6(cloud.covariances).numpy(), opacity=cloud.opacities.numpy(),
             H=H, W=W, bg=np.array([0.1, 0.3, 0.7], np.float32), **cam)
    gen = torch.Generator().manual_seed(seed + 1000)
    if color == "sh":
        d["shs"] = synthetic.random_sh(G, 3, sh_degree, seed=seed + 1).transpose(1, 2).contiguous().numpy()  # (G,n,3)
        d["sh_degree"] = sh_degree
    elif color == "precomp":
        d["colors_precomp"] = torch.rand(G, 3, generator=gen).numpy()
    if C:
        d["features"] = (0.5 + torch.randn(G, C, generator=gen) * 0.5).numpy()
    return d


def assert_close_with_flips(got, want, flip_bound, scale_vals, rtol=1e-4, atol=1e-5, what=""):
    """|got-want| <= rtol*|want| + atol, plus -- only where the oracle's own keep/skip decision was
    within its margin -- twice the blend weight that decision controls times the channel range."""
    got, want = np.asarray(got, np.float64), np.asarray(want, np.float64)
    tol = rtol * np.abs(want) + atol + 2.0 * np.asarray(flip_bound, np.float64) * scale_vals
    bad = np.abs(got - want) > tol
    assert not bad.any(), f"{what}: {bad.sum()} of {bad.size} outside tolerance, max err {np.abs(got - want).max():.3e}"


def rel_err(got, want):
    got, want = np.asarray(got, np.float64), np.asarray(want, np.float64)
    return np.abs(got - want).max() / max(np.abs(want).max(), 1e-30)


def init_by_name(module, seed: int = 0):
    """Deterministic, NAME-KEYED parameter values: two modules with the same parameter tree get identical
    weights regardless of construction order (used to compare our modules with the reference's without
    shipping weights).  Returns the sorted (name, shape) inventory."""
    import zlib
    inv = []
    with torch.no_grad():
        for name, p in sorted(module.named_parameters()):
            g = torch.Generator().manual_seed((zlib.crc32(name.encode()) + seed) % (2 ** 31))
            if p.dim() > 1:
                fan_in = p[0].numel()
                v = torch.randn(p.shape, generator=g) / math.sqrt(fan_in)
            elif "norm" in name.lower() and name.endswith("weight"):
                v = 1 + 0.1 * torch.randn(p.shape, generator=g)
            else:
                v = 0.1 * torch.randn(p.shape, generator=g)
            p.copy_(v.to(p.dtype))
            inv.append((name, tuple(p.shape)))
    return inv
