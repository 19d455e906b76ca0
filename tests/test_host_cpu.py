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


def test_harmonics_transform_equals_mask_then_rotate():
    """GaussianAdapter.harmonics_transform(c2w) @ raw == the adapter's own masking + per-degree SH rotation of the raw
    colour / feature blocks (gaussian_adapter.py:44-61, 104-105) and the identity on scales / rotation."""
    import torch
    from latentsplat_b200.model.encoder.common.gaussian_adapter import GaussianAdapter, GaussianAdapterCfg
    torch.manual_seed(0)
    ad = GaussianAdapter(GaussianAdapterCfg(0.5, 15.0, 4, 2), n_feature_channels=8).double()
    q = torch.nn.functional.normalize(torch.randn(2, 3, 4, dtype=torch.float64), dim=-1)
    from latentsplat_b200.model.encoder.common.gaussians import quaternion_to_matrix
    rot = quaternion_to_matrix(q)                                           # (2, 3, 3, 3) proper rotations
    raw = torch.randn(2, 3, 5, ad.d_in, dtype=torch.float64)               # 5 rays per view
    t = ad.harmonics_transform(rot)                                         # (2, 3, d_in, d_in)
    folded = torch.einsum("bvij,bvrj->bvri", t, raw)
    scales, rots, csh, fsh = raw.split((3, 4, 3 * ad.d_color_sh, 8 * ad.d_feature_sh), dim=-1)
    csh = csh.unflatten(-1, (3, ad.d_color_sh)) * ad.color_sh_mask
    fsh = fsh.unflatten(-1, (8, ad.d_feature_sh)) * ad.feature_sh_mask
    r = rot[:, :, None]                                                     # broadcast over rays
    want = torch.cat((scales, rots, ad._rotate(csh, r, 4).flatten(-2), ad._rotate(fsh, r, 2).flatten(-2)), dim=-1)
    torch.testing.assert_close(folded, want, rtol=1e-9, atol=1e-10)


def test_epipolar_image_index_matches_the_transposes_around_grid_sample():
    """EpipolarSampler.image_index(b, rays)[row] is the feature map the explicit transpose -> grid_sample -> transpose
    sequence samples for that row: with every feature map constant (= its flattened index) the sampled features say so."""
    import torch
    from latentsplat_b200.model.encoder.epipolar.epipolar_sampler import EpipolarSampler
    b, v, c, h, w = 2, 2, 3, 8, 8
    sampler = EpipolarSampler(v, 4)
    images = torch.arange(b * v, dtype=torch.float32).reshape(b, v, 1, 1, 1).expand(b, v, c, h, w).contiguous() + 1
    ex = torch.eye(4).repeat(b, v, 1, 1)
    ex[:, 1, 0, 3] = 0.2
    intr = torch.tensor([[0.9, 0, 0.5], [0, 0.9, 0.5], [0, 0, 1.0]]).repeat(b, v, 1, 1)
    near, far = torch.full((b, v), 1.0), torch.full((b, v), 20.0)
    s = sampler(images, ex, intr, near, far)
    idx = sampler.image_index(b, h * w).reshape(b, v, v - 1, h * w)
    feat = s.features[..., 0]                                               # (b, v, ov, r, s)
    inside = (s.xy_sample > 0.1).all(-1) & (s.xy_sample < 0.9).all(-1) & s.valid[..., None]
    assert inside.any()
    want = (idx[..., None].float() + 1).expand_as(feat)
    torch.testing.assert_close(feat[inside], want[inside])


def test_flat_gradients_follow_channels_last_parameters():
    import torch
    from latentsplat_b200.parallel import FlatGradients
    conv = torch.nn.Conv2d(4, 6, 3).to(memory_format=torch.channels_last)
    fg = FlatGradients(conv.parameters())
    assert conv.weight.grad.stride() == conv.weight.stride()
    conv(torch.randn(2, 4, 8, 8)).sum().backward()
    assert float(fg.flat.abs().sum()) > 0 and fg.flat.numel() == sum(p.numel() for p in conv.parameters())


def test_math_mode_switch_turns_the_contraction_kernels_off_and_on():
    """precision choices are one documented switch (latentsplat_b200.precision), not scattered module globals."""
    import torch
    from latentsplat_b200 import conv, fmha, gemm, precision
    try:
        precision.set_math_mode("fp32")
        assert not (gemm.enabled or conv.ENABLED or fmha.ENABLED) and not torch.backends.cudnn.allow_tf32
        assert precision.math_mode() == "fp32"
    finally:
        precision.set_math_mode("tf32")
    assert gemm.enabled and conv.ENABLED and fmha.ENABLED and precision.math_mode() == "tf32"
    import pytest
    with pytest.raises(ValueError):
        precision.set_math_mode("bf16")


def test_closed_form_camera_inverses_match_lapack():
    """geometry/inverse.py (adjugate formulas, no cuSOLVER) against torch.linalg.inv in float64, incl. autograd."""
    import torch
    from latentsplat_b200 import synthetic
    from latentsplat_b200.geometry.inverse import inv2x2, inv3x3, inv_affine4x4
    g = torch.Generator().manual_seed(0)
    pose = torch.stack([synthetic.pose(0.3 * i, -4.0 + i, 0.05 * i, 0.02) for i in range(5)]).double()
    pose[:, :3, :3] *= 1.7                                       # a scaled rigid transform is still affine
    k3 = synthetic.intrinsics(0.9)[None].repeat(5, 1, 1).double() + 0.01 * torch.randn(5, 3, 3, generator=g).double()
    for fn, m in ((inv_affine4x4, pose), (inv3x3, k3), (inv2x2, k3[:, :2, :2])):
        m = m.clone().requires_grad_(True)
        ours, ref = fn(m), torch.linalg.inv(m)
        torch.testing.assert_close(ours, ref, rtol=1e-10, atol=1e-12)
        w = torch.randn(ours.shape, generator=g).double()
        (ga,) = torch.autograd.grad((ours * w).sum(), m)
        (gb,) = torch.autograd.grad((torch.linalg.inv(m) * w).sum(), m)
        rows = 3 if fn is inv_affine4x4 else m.shape[-2]        # the (0, 0, 0, 1) row of a pose is a constant, not a variable
        torch.testing.assert_close(ga[:, :rows], gb[:, :rows], rtol=1e-8, atol=1e-10)
    assert torch.equal(inv_affine4x4(pose.float())[:, 3], torch.tensor([0.0, 0.0, 0.0, 1.0]).expand(5, 4))
