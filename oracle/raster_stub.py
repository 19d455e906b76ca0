"""A CPU `diff_gaussian_rasterization` look-alike backed by the oracle (TEST INFRASTRUCTURE; lives in oracle/).

It lets the reference's own Python (render_cuda, DecoderSplattingCUDA) run unmodified on the CPU
box to generate golden fixtures, and gives the tests a per-view reference with autograd.
Never imported by latentsplat_b200/.
"""
from __future__ import annotations

from typing import NamedTuple

import numpy as np
import torch
from torch import Tensor, nn

from oracle import oracle


class OracleSettings(NamedTuple):
    image_height: int
    image_width: int
    tanfovx: float
    tanfovy: float
    bg: Tensor
    scale_modifier: float
    viewmatrix: Tensor
    projmatrix: Tensor
    sh_degree: int
    campos: Tensor
    prefiltered: bool
    debug: bool


def _np(t):
    return None if t is None else t.detach().cpu().numpy()


class _OracleFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, means3D, means2D, opacities, cov3D, shs, colors_precomp, features, rs, prec):
        r = oracle.forward(means3D=_np(means3D), cov3D=_np(cov3D), opacity=_np(opacities).reshape(-1),
                           viewmatrix=_np(rs.viewmatrix), projmatrix=_np(rs.projmatrix), campos=_np(rs.campos),
                           tanfovx=float(rs.tanfovx), tanfovy=float(rs.tanfovy), H=rs.image_height,
                           W=rs.image_width, bg=_np(rs.bg), shs=_np(shs), colors_precomp=_np(colors_precomp),
                           features=_np(features), sh_degree=rs.sh_degree, prec=prec, n_threads=OracleGaussianRasterizer.n_threads)
        ctx.r = r
        ctx.shapes = (opacities.shape,)
        dt = means3D.dtype
        t = lambda a: None if a is None else torch.from_numpy(np.asarray(a)).to(dt)
        outs = (t(r.out_color), t(r.out_feature), t(r.out_alpha)[None], t(r.out_depth)[None],
                torch.from_numpy(r.radii.copy()))
        ctx.mark_non_differentiable(outs[4])
        return outs

    @staticmethod
    def backward(ctx, g_color, g_feature, g_alpha, g_depth, _):
        r = ctx.r
        g = oracle.backward(r, dL_dcolor=_np(g_color), dL_dfeature=_np(g_feature),
                            dL_dalpha=None if g_alpha is None else _np(g_alpha)[0],
                            dL_ddepth=None if g_depth is None else _np(g_depth)[0],
                            n_threads=1 if OracleGaussianRasterizer.n_threads == 0 and not OracleGaussianRasterizer.parallel_backward
                            else OracleGaussianRasterizer.n_threads)
        dt = torch.float64 if r.prec == "f64" else torch.float32
        t = lambda a: None if a is None else torch.from_numpy(np.asarray(a)).to(dt)
        d_means2D = torch.cat([t(g["dL_dmeans2D"]), torch.zeros(r.G, 1, dtype=dt)], dim=1)
        d_shs = t(g["dL_dshs"]) if r.shs is not None else None
        d_colors = t(g["dL_dcolors"]) if r.colors_precomp is not None else None
        return (t(g["dL_dmeans3D"]), d_means2D, t(g["dL_dopacity"]).reshape(ctx.shapes[0]), t(g["dL_dcov3D"]), d_shs,
                d_colors, t(g["dL_dfeatures"]), None, None)


class OracleGaussianRasterizer(nn.Module):
    """Same call signature as the rasterizer object built at cuda_splatting.py:146-158."""
    prec = "f32"
    n_threads = 0              # forward OpenMP threads (0 = all cores)
    parallel_backward = False  # tests keep the backward single-threaded (deterministic sums); bench.py turns it on

    def __init__(self, raster_settings):
        super().__init__()
        self.raster_settings = raster_settings

    def forward(self, means3D, means2D, opacities, shs=None, colors_precomp=None, features=None, scales=None,
                rotations=None, cov3D_precomp=None):
        assert cov3D_precomp is not None
        prec = "f64" if means3D.dtype == torch.float64 else self.prec
        return _OracleFn.apply(means3D, means2D, opacities, cov3D_precomp, shs, colors_precomp, features,
                               self.raster_settings, prec)
