"""GPU: the fused encoder tail (ls_gaussian_head_*, include/ls_ghead.h) against the explicit torch sequence it replaces
(DepthPredictorMonocular -> xy offsets -> opacity mapping -> GaussianAdapter, our parameter-compatible modules, which the CPU
tests pin to the reference's own modules through tests/golden/encoder.npz).  Forward values and the gradients w.r.t. both
Linear-head outputs; fp32 both ways, so the tolerance is rounding-level (1e-4 relative + RMS-scaled floor)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _close(a, b, what, rtol=2e-4):
    a, b = a.double(), b.double()
    rms = b.square().mean().sqrt().item() + 1e-30
    err = (a - b).abs()
    bad = err > rtol * (b.abs() + rms)
    assert bad.float().mean().item() <= 1e-3, f"{what}: {int(bad.sum())} of {bad.numel()} off, max err {err.max().item():.3e} (rms {rms:.3e})"


def _reference_tail(dlog, raw, cams, h, w, spp, deterministic, gpp, adapter, exponent):
    """The tail of EncoderEpipolar.forward as explicit torch ops (same code path as the non-fused branch)."""
    from latentsplat_b200.geometry.projection import sample_image_grid
    from latentsplat_b200.misc.discrete_probability_distribution import gather_discrete_topk, sample_discrete_distribution
    from latentsplat_b200.model.encoder.epipolar.conversions import relative_disparity_to_depth
    ext, intr, near, far = cams
    b, v, r, _ = dlog.shape
    x = dlog.unflatten(-1, (32, 1, 2)).movedim(-1, 0).transpose(-1, -2)
    pdf, offset = x[0].softmax(dim=-1), x[1].sigmoid()
    index, pdf_i = gather_discrete_topk(pdf, spp) if deterministic else sample_discrete_distribution(pdf, spp)
    offset = offset.gather(-1, index)
    depth = relative_disparity_to_depth((index + offset) / 32, near[..., None, None, None], far[..., None, None, None])
    opacity = 0.5 * (1 - (1 - pdf_i) ** exponent + pdf_i ** (1 / exponent)) / gpp
    g = raw.unflatten(-1, (1, -1))
    xy, _ = sample_image_grid((h, w), dlog.device)
    xy = xy.reshape(h * w, 1, 2) + (g[..., :2].sigmoid() - 0.5) * torch.tensor((1 / w, 1 / h), device=dlog.device)
    out = adapter(ext[:, :, None, None, None], intr[:, :, None, None, None], xy[..., None, :], depth, opacity,
                  g[..., None, 2:], (h, w), harmonics_ready=True)
    return (out.means.flatten(1, 4), out.covariances.flatten(1, 4), out.opacities.flatten(1, 4),
            out.color_harmonics.flatten(1, 4).flatten(-2), out.feature_harmonics.flatten(1, 4).flatten(-2), index)


@pytest.mark.parametrize("deterministic,exponent", [(True, 1.0), (False, 1.0), (False, 2.0 ** 0.5)])
def test_fused_gaussian_head_matches_explicit_sequence(cuda, deterministic, exponent):
    from latentsplat_b200 import gaussian_head, synthetic
    from latentsplat_b200.model.encoder.common.gaussian_adapter import GaussianAdapter, GaussianAdapterCfg
    b, v, h, w, C = 2, 2, 16, 24, 8
    r, spp, gpp = h * w, (1 if deterministic else 3), 3
    adapter = GaussianAdapter(GaussianAdapterCfg(0.5, 15.0, 4, 2), C).to(cuda)
    g = torch.Generator(cuda).manual_seed(3)
    dlog = (2 * torch.randn(b, v, r, 64, device=cuda, generator=g)).requires_grad_(True)
    raw = torch.randn(b, v, r, 9 + 75 + 9 * C, device=cuda, generator=g).requires_grad_(True)
    ext = torch.stack([torch.stack([synthetic.pose(0.3 * i + 0.1 * j, 7.0 * j - 3.0 * i) for j in range(v)]) for i in range(b)]).to(cuda)
    intr = synthetic.intrinsics(0.86)[None, None].repeat(b, v, 1, 1).to(cuda)
    intr[1, 0, 0, 0], intr[1, 0, 0, 1] = 1.1, 0.05                                  # a skewed, anisotropic camera as well
    near = torch.full((b, v), 0.7, device=cuda) + 0.1 * torch.arange(v, device=cuda)
    far = torch.full((b, v), 60.0, device=cuda)
    wts = [torch.randn(s, device=cuda, generator=g) for s in ((b, v * r * spp, 3), (b, v * r * spp, 3, 3), (b, v * r * spp),
                                                               (b, v * r * spp, 75), (b, v * r * spp, 9 * C))]

    torch.manual_seed(11)
    u = None if deterministic else torch.rand((b, v, r, 1, spp), device=cuda).reshape(b, v, r, spp)
    got = gaussian_head.gaussian_head(dlog, raw, u, ext, intr, near, far, (h, w), spp, 75, 9 * C, 0.5, 15.0, exponent, gpp)
    loss = sum((o * wt).sum() for o, wt in zip(got[:5], wts))
    gd, gr = torch.autograd.grad(loss, (dlog, raw))

    torch.manual_seed(11)                                                           # the explicit path draws the same numbers
    ref = _reference_tail(dlog, raw, (ext, intr, near, far), h, w, spp, deterministic, gpp, adapter, exponent)
    loss_r = sum((o * wt).sum() for o, wt in zip(ref[:5], wts))
    rd, rr = torch.autograd.grad(loss_r, (dlog, raw))

    same = (got[5].view(-1) == ref[5].reshape(-1).to(torch.int32))
    assert same.float().mean().item() >= 0.999, "sampled depth buckets differ"
    keep = same.view(b, v * r * spp)                                                # compare Gaussians whose bucket agrees
    for a, e, name in zip(got[:5], ref[:5], ("means", "covariances", "opacities", "colour SH", "feature SH")):
        _close(a[keep], e[keep], name)
    ray_ok = same.view(b, v, r, spp).all(dim=-1)
    _close(gd[ray_ok], rd[ray_ok], "d depth logits")
    _close(gr[ray_ok], rr[ray_ok], "d raw Gaussian parameters")


def test_fused_reparameterised_sample_matches_torch(cuda):
    """DiagonalGaussianDistribution.sample() on the packed params (ls_reparam_*): same draws (torch RNG, same call), same values and
    gradients as mean + exp(0.5 clamp(logvar)) * eps, including the clamp's closed-interval gradient."""
    import torch
    from latentsplat_b200.model.diagonal_gaussian_distribution import DiagonalGaussianDistribution
    g = torch.Generator(cuda).manual_seed(3)
    params = torch.randn(3, 1000, 8, 9, device=cuda, generator=g) * 12.0          # logvars beyond both clamp bounds
    params[0, 0, 4, 0], params[0, 1, 4, 0] = -30.0, 20.0                            # exactly on the bounds: gradient passes
    w = torch.randn(3, 1000, 4, 9, device=cuda, generator=g)
    res = []
    for fused in (True, False):
        p = params.clone().requires_grad_(True)
        d = DiagonalGaussianDistribution(params=p if fused else None, mean=None if fused else p[:, :, :4], logvar=None if fused else p[:, :, 4:],
                                         dim=2)
        torch.manual_seed(77)
        s = d.sample()
        (s * w).sum().backward()
        res.append((s.detach(), p.grad.clone()))
    assert torch.allclose(res[0][0], res[1][0], rtol=1e-5, atol=1e-5)
    assert torch.allclose(res[0][1], res[1][1], rtol=1e-5, atol=1e-5)
    assert res[0][1][0, 0, 4, 0] != 0 and res[0][1][0, 1, 4, 0] != 0
