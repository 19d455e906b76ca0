"""CPU: the loss heads (latentsplat_b200/loss) against goldens computed by the REFERENCE's own
src/loss modules (tests/golden/make_golden.py::loss_goldens), and the restated LPIPS-VGG against the published definition
assembled from torchvision's VGG-16 layers."""
import importlib.util
from pathlib import Path

import numpy as np
import pytest
import torch

GOLD = Path(__file__).parent / "golden"


def _mg():
    spec = importlib.util.spec_from_file_location("make_golden", GOLD / "make_golden.py")
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def test_loss_groups_match_reference_goldens():
    from latentsplat_b200.loss import (LossDepthCfg, LossDiscriminatorCfg, LossGeneratorCfg, LossGroupCfg, LossKlCfg, LossL1Cfg,
                                       LossMseCfg, get_loss_group)
    from latentsplat_b200.model.diagonal_gaussian_distribution import DiagonalGaussianDistribution
    from latentsplat_b200.model.types import GroundTruth, Prediction
    g = np.load(GOLD / "losses.npz")
    x = _mg().loss_inputs()
    W = x["W"].clone().requires_grad_(True)
    image = torch.einsum("oc,bvchw->bvohw", W, x["x"])
    logits_fake = torch.einsum("oc,bvchw->bvohw", x["Wd"], image)[..., ::4, ::4]
    pred = Prediction(image=image, posterior=DiagonalGaussianDistribution(x["mean"], x["logvar"], dim=2), depth=x["depth"],
                      logits_fake=logits_fake, logits_real=x["logits_real"])
    gt = GroundTruth(image=x["gt_image"], near=x["near"], far=x["far"])
    checked = 0
    for tag, depth_cfg in (("plain", LossDepthCfg(weight=0.25)),
                           ("bilateral2", LossDepthCfg(weight=0.25, sigma_image=10.0, use_second_derivative=True))):
        for disc in ("hinge", "vanilla"):
            cfg = LossGroupCfg(nll=[LossMseCfg(weight=10.0), LossL1Cfg(weight=1.0), LossKlCfg(weight=1e-3), depth_cfg],
                               generator=LossGeneratorCfg(weight=0.5, apply_after_step=3),
                               discriminator=LossDiscriminatorCfg(weight=2.0, loss=disc, apply_after_step=3))
            group = get_loss_group("target", cfg)
            assert group.is_active(0) and not group.is_generator_loss_active(0) and group.is_generator_loss_active(3)
            for step in (0, 7):
                key = f"{tag}_{disc}_s{step}"
                total, d = group.forward_generator(pred, gt, step, last_layer_weights=W)
                np.testing.assert_allclose(total.detach().numpy(), g[f"{key}_gen_total"], rtol=1e-5)
                assert sorted(f"{key}_gen_{k}_u" for k in d) == sorted(k for k in g.files if k.startswith(f"{key}_gen_") and k.endswith("_u"))
                for k, v in d.items():
                    np.testing.assert_allclose(v.unweighted.detach().numpy(), g[f"{key}_gen_{k}_u"], rtol=1e-5, atol=1e-7, err_msg=k)
                    np.testing.assert_allclose(v.weighted.detach().numpy(), g[f"{key}_gen_{k}_w"], rtol=1e-5, atol=1e-7, err_msg=k)
                    checked += 2
                if step >= 3:
                    total, d = group.forward_discriminator(pred, gt, step)
                    np.testing.assert_allclose(total.detach().numpy(), g[f"{key}_dis_total"], rtol=1e-5)
                    for k, v in d.items():
                        np.testing.assert_allclose(v.weighted.detach().numpy(), g[f"{key}_dis_{k}_w"], rtol=1e-5, err_msg=k)
                        checked += 1
    assert checked >= 60
    # the adaptive weight is in (0, 1] and detached: the generator term's weighted value carries no graph through it
    total, d = group.forward_generator(pred, gt, 7, last_layer_weights=W)
    (gw,) = torch.autograd.grad(total, W)
    assert torch.isfinite(gw).all()


def test_lpips_vgg_matches_the_published_definition():
    """Same weights in torchvision's VGG-16 + the published LPIPS formula written out here == latentsplat_b200.loss.LpipsVgg."""
    torchvision = pytest.importorskip("torchvision")
    import torch.nn.functional as F
    from latentsplat_b200.loss import LossLpips, LossLpipsCfg, LpipsVgg
    from latentsplat_b200.model.types import GroundTruth, Prediction
    torch.manual_seed(1)
    vgg = torchvision.models.vgg16(weights=None).features.eval()
    lins = [torch.rand(1, c, 1, 1) for c in (64, 128, 256, 512, 512)]
    state = {f"features.{i}.{n}": getattr(m, n).detach() for i, m in enumerate(vgg) if isinstance(m, torch.nn.Conv2d) for n in ("weight", "bias")}
    state.update({f"lin{k}.model.1.weight": w for k, w in enumerate(lins)})
    import tempfile
    with tempfile.TemporaryDirectory() as tmp:
        path = Path(tmp) / "lpips_vgg.pth"
        torch.save(state, path)
        ours = LpipsVgg(weights=str(path))
    assert not any(p.requires_grad for p in ours.parameters()) and len(ours.state_dict()) == 31      # 13 x (w, b) + 5 lins; buffers off
    a, b = torch.rand(3, 3, 48, 40), torch.rand(3, 3, 48, 40)

    def published(x0, x1):
        shift, scale = torch.tensor([-0.030, -0.088, -0.188]).view(1, 3, 1, 1), torch.tensor([0.458, 0.448, 0.450]).view(1, 3, 1, 1)
        f0, f1 = (2 * x0 - 1 - shift) / scale, (2 * x1 - 1 - shift) / scale
        taps, total = (3, 8, 15, 22, 29), 0.0                 # relu1_2, relu2_2, relu3_3, relu4_3, relu5_3
        k = 0
        for i, layer in enumerate(vgg):
            f0, f1 = layer(f0), layer(f1)
            if i in taps:
                n0 = f0 / (f0.pow(2).sum(1, keepdim=True).sqrt() + 1e-10)
                n1 = f1 / (f1.pow(2).sum(1, keepdim=True).sqrt() + 1e-10)
                total = total + F.conv2d((n0 - n1) ** 2, lins[k]).mean(dim=(2, 3), keepdim=True)
                k += 1
            if i == 29:
                break
        return total.mean()

    with torch.no_grad():
        want, got = published(a, b), ours(a, b)
    torch.testing.assert_close(got, want, rtol=2e-5, atol=1e-7)
    loss = LossLpips(LossLpipsCfg(weight=0.05), lpips=ours)
    v = loss(Prediction(image=a[None]), GroundTruth(image=b[None]), 0)
    torch.testing.assert_close(v.weighted, 0.05 * want, rtol=2e-5, atol=1e-8)
    with pytest.warns(UserWarning, match="RANDOMLY initialised"):
        LpipsVgg()
