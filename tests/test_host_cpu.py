"""CPU: host logic mirrored from the reference (camera maths, carriers, distribution) against golden
vectors produced by the reference's own functions (tests/golden/make_golden.py)."""
from pathlib import Path

import numpy as np
import pytest
import torch

GOLD = Path(__file__).parent / "golden"


def test_fov_projection_and_view_matrices_match_reference():
    from latentsplat_b200.model.decoder.cuda_splatting import _camera_matrices, get_fov, get_projection_matrix
    g = np.load(GOLD / "camera.npz")
    t = lambda k: torch.from_numpy(g[k])
    fov = get_fov(t("intrinsics"))
    np.testing.assert_allclose(fov.numpy(), g["fov"], rtol=1e-6)
    proj = get_projection_matrix(t("near"), t("far"), fov[:, 0], fov[:, 1])
    np.testing.assert_allclose(proj.numpy(), g["projection"], rtol=1e-6, atol=1e-7)
    view_t, full_t = _camera_matrices(t("extrinsics"), t("near"), t("far"), fov[:, 0], fov[:, 1])
    np.testing.assert_allclose(view_t.numpy(), g["view_t"], rtol=1e-6, atol=1e-7)
    np.testing.assert_allclose(full_t.numpy(), g["full_t"], rtol=1e-6, atol=1e-6)


def test_upper_triangle_order():
    from latentsplat_b200.model.decoder.cuda_splatting import _upper_triangle
    c = torch.arange(9.0).reshape(3, 3)
    i, j = torch.triu_indices(3, 3)
    assert torch.equal(_upper_triangle(c), c[i, j])


def test_diagonal_gaussian_distribution_semantics():
    from latentsplat_b200.model.diagonal_gaussian_distribution import DiagonalGaussianDistribution as D
    mean, logvar = torch.randn(2, 4, 3), torch.tensor(50.0).expand(2, 4, 3)
    d = D(mean, logvar)
    assert float(d.logvar.max()) == 20.0  # clamped to (-30, 20)
    assert torch.equal(d.params, torch.cat((mean, d.logvar), dim=0))
    d0 = D(mean)  # zero variance
    assert d0.sample() is mean and float(d0.kl().abs().max()) == 0 and d0.logvar is None
    p = torch.randn(2, 8, 9)
    dp = D(params=p, dim=-2)
    assert torch.equal(dp.mean, p[:, :4]) and torch.equal(dp.logvar, p[:, 4:].clamp(-30, 20))
    torch.manual_seed(0)
    s = dp.sample()
    torch.manual_seed(0)
    assert torch.equal(s, dp.mean + dp.std * torch.randn_like(dp.mean))
    np.testing.assert_allclose(dp.kl().numpy(), 0.5 * (dp.mean ** 2 + dp.var - 1 - dp.logvar).numpy())
    with pytest.raises(AssertionError):
        D(mean, logvar, params=p)


def test_variational_gaussians_sample_mode_flatten():
    from latentsplat_b200.model.diagonal_gaussian_distribution import DiagonalGaussianDistribution as D
    from latentsplat_b200.model.types import Gaussians, VariationalGaussians
    p = torch.randn(2, 5, 8, 9)
    vg = VariationalGaussians(torch.zeros(2, 5, 3), torch.zeros(2, 5, 3, 3), torch.ones(2, 5), None, D(params=p, dim=-2))
    assert isinstance(vg.mode(), Gaussians) and vg.mode().feature_harmonics.shape == (2, 5, 4, 9)
    assert torch.equal(vg.flatten().feature_harmonics, p)
    assert vg.sample().feature_harmonics.shape == (2, 5, 4, 9)


def test_synthetic_generator_is_deterministic_and_in_frustum():
    from latentsplat_b200 import synthetic
    a = synthetic.random_gaussians(1000, seed=3)
    b = synthetic.random_gaussians(1000, seed=3)
    assert torch.equal(a.means, b.means) and torch.equal(a.covariances, b.covariances)
    assert (a.means[:, 2] >= 1.0 - 1e-5).all() and (a.means[:, 2] <= 100.0 + 1e-3).all()
    ev = torch.linalg.eigvalsh(a.covariances.double())
    assert (ev > 0).all()
    assert synthetic.sh_mask(2).tolist() == pytest.approx([1.0] + [0.025] * 3 + [0.00625] * 5)
