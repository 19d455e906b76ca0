"""CPU: latentsplat_b200.trainer.TrainStep (sync-free generator + discriminator optimisation) against a
plain re-statement of the reference's control flow (model_wrapper.py:424-489: zero_grad / manual_backward / clip_gradients(0.5,
"norm") / step, discriminator on the detached prediction) with stock torch optimisers; and the device-side NaN guard."""
import copy

import torch
from torch import nn

from latentsplat_b200.loss import (LossDiscriminatorCfg, LossGeneratorCfg, LossGroupCfg, LossL1Cfg, LossMseCfg, get_loss_group)
from latentsplat_b200.model.types import GroundTruth, Prediction
from latentsplat_b200.trainer import OptimizerCfg, TrainStep


class _Gen(nn.Module):
    def __init__(self):
        super().__init__()
        self.body = nn.Conv2d(3, 8, 3, padding=1)
        self.render_head = nn.Conv2d(8, 3, 1)
        self.last = nn.Conv2d(8, 3, 3, padding=1)              # "conv_out": the layer the adaptive weight looks at

    def forward(self, x):
        h = torch.tanh(self.body(x.flatten(0, 1)))
        b, v = x.shape[:2]
        return self.render_head(h).unflatten(0, (b, v)), self.last(h).unflatten(0, (b, v))


def _setup(seed=0):
    torch.manual_seed(seed)
    gen, disc = _Gen(), nn.Sequential(nn.Conv2d(3, 4, 4, stride=2, padding=1), nn.LeakyReLU(0.2), nn.Conv2d(4, 1, 4, stride=2, padding=1))
    g = torch.Generator().manual_seed(seed + 1)
    batches = [{"context": {"image": torch.rand(2, 2, 3, 16, 16, generator=g)},
                "target": {"image": torch.rand(2, 2, 3, 16, 16, generator=g), "near": torch.ones(2, 2), "far": torch.full((2, 2), 9.0)}}
               for _ in range(3)]
    return gen, disc, batches


def _groups():
    render = get_loss_group("target/render/image", LossGroupCfg(nll=[LossMseCfg(weight=10)]))
    combined = get_loss_group("target/combined", LossGroupCfg(nll=[LossL1Cfg()], generator=LossGeneratorCfg(weight=0.5, apply_after_step=1),
                                                              discriminator=LossDiscriminatorCfg(apply_after_step=1)))
    return render, combined


def test_train_step_equals_reference_control_flow():
    gen, disc, batches = _setup()
    ref_gen, ref_disc = copy.deepcopy(gen), copy.deepcopy(disc)
    render, combined = _groups()

    def forward_fn(batch):
        r, c = gen(batch["context"]["image"])
        return Prediction(image=r), Prediction(image=c)

    step = TrainStep(forward_fn, gen.parameters(), disc, render, combined, gen.last.weight, OptimizerCfg(lr=1e-2), OptimizerCfg(lr=2e-2))
    logs = [step(b, i) for i, b in enumerate(batches)]
    assert "discriminator/total" not in logs[0] and "discriminator/total" in logs[1]          # apply_after_step = 1
    assert all(bool(l["generator/finite"]) for l in logs)

    # --- the reference's flow, spelled out with stock optimisers ---
    g_opt = torch.optim.Adam(ref_gen.parameters(), lr=1e-2)
    d_opt = torch.optim.Adam(ref_disc.parameters(), lr=2e-2)
    rr, rc = _groups()
    for i, batch in enumerate(batches):
        r, c = ref_gen(batch["context"]["image"])
        gt = GroundTruth(image=batch["target"]["image"], near=batch["target"]["near"], far=batch["target"]["far"])
        rp, cp = Prediction(image=r), Prediction(image=c)
        if rc.is_generator_loss_active(i):
            for p in ref_disc.parameters():
                p.requires_grad_(False)
            cp.logits_fake = ref_disc(c.flatten(0, 1)).unflatten(0, c.shape[:2])
        loss = rr.forward_generator(rp, gt, i, ref_gen.last.weight)[0] + rc.forward_generator(cp, gt, i, ref_gen.last.weight)[0]
        g_opt.zero_grad()
        loss.backward()
        torch.nn.utils.clip_grad_norm_(ref_gen.parameters(), 0.5)
        g_opt.step()
        for p in ref_disc.parameters():
            p.requires_grad_(True)
        torch.testing.assert_close(logs[i]["generator/total"], loss.detach(), rtol=1e-5, atol=1e-7)
        if rc.is_discriminator_loss_active(i):
            pred = Prediction(image=c.detach())
            pred.logits_fake = ref_disc(pred.image.flatten(0, 1)).unflatten(0, c.shape[:2])
            pred.logits_real = ref_disc(batch["target"]["image"].flatten(0, 1)).unflatten(0, c.shape[:2])
            dl = rc.forward_discriminator(pred, gt, i)[0]
            d_opt.zero_grad()
            dl.backward()
            torch.nn.utils.clip_grad_norm_(ref_disc.parameters(), 0.5)
            d_opt.step()
            torch.testing.assert_close(logs[i]["discriminator/total"], dl.detach(), rtol=1e-5, atol=1e-7)
    for (n, a), b in zip(gen.named_parameters(), ref_gen.parameters()):
        torch.testing.assert_close(a, b, rtol=1e-5, atol=1e-7, msg=f"generator {n}")
    for (n, a), b in zip(disc.named_parameters(), ref_disc.parameters()):
        torch.testing.assert_close(a, b, rtol=1e-5, atol=1e-7, msg=f"discriminator {n}")


def test_non_finite_loss_leaves_the_parameters_alone_without_a_host_sync():
    gen, disc, batches = _setup(3)
    render, combined = _groups()

    def forward_fn(batch):
        r, c = gen(batch["context"]["image"])
        return Prediction(image=r), Prediction(image=c)

    step = TrainStep(forward_fn, gen.parameters(), disc, render, combined, gen.last.weight, OptimizerCfg(lr=1e-2), OptimizerCfg(lr=1e-2))
    step(batches[0], 5)
    before = [p.detach().clone() for p in list(gen.parameters()) + list(disc.parameters())]
    bad = copy.deepcopy(batches[1])
    bad["target"]["image"][0, 0, 0, 0, 0] = float("nan")
    log = step(bad, 6)
    assert not bool(log["generator/finite"]) and not bool(log["discriminator/finite"])
    for a, b in zip(before, list(gen.parameters()) + list(disc.parameters())):
        assert torch.equal(a, b.detach())
    good = step(batches[2], 7)                                  # and training goes on
    assert bool(good["generator/finite"]) and torch.isfinite(good["generator/grad_norm"])
    assert any(not torch.equal(a, b.detach()) for a, b in zip(before, gen.parameters()))
