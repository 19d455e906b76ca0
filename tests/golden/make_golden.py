"""Generates the committed golden fixtures in tests/golden/ by importing the REFERENCE's own
Python (needs /root/reference, i.e. the build container; the GPU box only reads the .npz).

  python tests/golden/make_golden.py

Fixtures
  sh_eval.npz      /root/reference/src/misc/sh_utils.py:42-97 `eval_sh` (float64) for degrees 0..4
                   -> pins the SH basis of the oracle and of the CUDA preprocess kernel.
  camera.npz       `get_fov` (src/geometry/projection.py:233-247) and `get_projection_matrix`
                   (src/model/decoder/cuda_splatting.py:19-46), plus the transposed view / full
                   projection matrices exactly as render_cuda builds them (:111-118).
  render_cuda.npz  the reference's `render_cuda` (cuda_splatting.py:56-167) and
                   `DecoderSplattingCUDA.forward` (decoder_splatting_cuda.py:58-91) run UNMODIFIED on the
                   CPU with `diff_gaussian_rasterization` provided by the CPU oracle (the rasterizer's own
                   source is not in the reference).  Pins the host logic around the
                   rasterizer: 1/near rescale, SH layouts, feature eval_sh + 0.5, matrix conventions,
                   per-view repeat, posterior construction.
Inputs are regenerated from seeds by latentsplat_b200.synthetic, only outputs are stored.
"""
from __future__ import annotations

import importlib.abc
import importlib.machinery
import sys
import types
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[2]
REF = Path("/root/reference")
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

from oracle import oracle  # noqa: E402
from latentsplat_b200 import synthetic  # noqa: E402


def _stub(name: str, **attrs):
    m = types.ModuleType(name)
    m.__spec__ = importlib.machinery.ModuleSpec(name, None)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


MISSING = ("dacite", "hydra", "omegaconf", "pytorch_lightning", "lightning", "lightning_fabric", "diffusers", "timm",
           "lpips", "e3nn", "DISTS_pytorch", "skimage", "plyfile", "moviepy", "matplotlib", "colorspacious",
           "colorama", "svg", "cv2", "wandb", "torchmetrics")


class _DummyMeta(type):
    def __getattr__(cls, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return cls


class _Dummy(metaclass=_DummyMeta):
    """Stands in for any class/function of a package that is absent in this container."""
    def __init__(self, *a, **k):
        pass

    def __call__(self, *a, **k):
        return self

    def __getattr__(self, name):
        return _Dummy()

    def __class_getitem__(cls, item):
        return cls

    def __mro_entries__(self, bases):
        return (_Dummy,)


class _StubFinder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    def find_spec(self, fullname, path, target=None):
        if fullname.split(".")[0] in MISSING:
            return importlib.machinery.ModuleSpec(fullname, self, is_package=True)
        return None

    def create_module(self, spec):
        m = types.ModuleType(spec.name)
        m.__path__ = []
        m.__getattr__ = lambda name: _Dummy if not name.startswith("__") else (_ for _ in ()).throw(AttributeError(name))
        return m

    def exec_module(self, module):
        pass


def install_reference():
    """Make `src.*` of the reference importable on this CPU box (stubs for absent third-party packages)."""
    if not REF.exists():
        raise SystemExit("/root/reference is not present: fixtures can only be generated in the build container")
    sys.meta_path.insert(0, _StubFinder())
    # the rasterizer: CPU oracle behind the reference's own call signature
    from oracle.raster_stub import OracleGaussianRasterizer, OracleSettings
    _stub("diff_gaussian_rasterization", GaussianRasterizationSettings=OracleSettings,
          GaussianRasterizer=OracleGaussianRasterizer)
    sys.path.insert(0, str(REF))


def scene_inputs(seed: int, G: int, b: int, v: int, C: int = 4, color_deg: int = 4, feat_deg: int = 2):
    """Inputs of DecoderSplattingCUDA.forward for b scenes x v target views (shared by the tests)."""
    ctx = synthetic.pose()
    means, covs, opacs, csh, fsh = [], [], [], [], []
    for i in range(b):
        cloud = synthetic.random_gaussians(G, seed=seed + i, extrinsics=ctx, near=1.0, far=20.0, width=64,
                                           s_px=(0.7, 4.0))
        means.append(cloud.means * 2.5)  # scene not in units of near, so that the 1/near rescale matters
        covs.append(cloud.covariances * 2.5 ** 2)
        opacs.append(cloud.opacities)
        csh.append(synthetic.random_sh(G, 3, color_deg, seed=seed + 100 + i) * 0.5)
        fsh.append(synthetic.random_sh(G, C, feat_deg, seed=seed + 200 + i))
    extr = synthetic.target_poses(v)[None].repeat(b, 1, 1, 1).clone()
    extr[..., :3, 3] *= 2.5
    intr = synthetic.intrinsics(0.9)[None, None].repeat(b, v, 1, 1).clone()
    intr[..., 1, 1] = 1.1
    near = torch.full((b, v), 2.5)
    far = torch.full((b, v), 50.0)
    return dict(means=torch.stack(means), covariances=torch.stack(covs), opacities=torch.stack(opacs),
                color_harmonics=torch.stack(csh), feature_harmonics=torch.stack(fsh), extrinsics=extr,
                intrinsics=intr, near=near, far=far)


RENDER_CFG = dict(seed=4321, G=1500, b=2, v=2, H=48, W=64)


def encoder_cfgs():
    """Small-but-complete configuration of the epipolar encoder (reference dims where the code hard-codes them)."""
    sa = dict(patch_size=4, num_octaves=10, num_layers=2, num_heads=4, d_token=64, d_dot=32, d_mlp=128)
    et = dict(num_octaves=10, num_layers=2, num_heads=4, num_samples=32, d_dot=32, d_mlp=128, downscale=4)
    ga = dict(gaussian_scale_min=0.5, gaussian_scale_max=15.0, color_sh_degree=4, feature_sh_degree=2)
    enc = dict(name="epipolar", d_backbone=96, d_feature=64, num_monocular_samples=32, num_surfaces=1,
               predict_opacity=False, near_disparity=3.0, apply_bounds_shim=True, gaussians_per_pixel=3,
               use_epipolar_transformer=True, use_transmittance=False)
    return sa, et, ga, enc


def encoder_context(b=1, v=2, hw=32, seed=5):
    gen = torch.Generator().manual_seed(seed)
    extr = torch.stack([synthetic.pose(0.0, 0.0), synthetic.pose(0.6, -6.0, 0.05, 0.02)])[None].repeat(b, 1, 1, 1)
    intr = synthetic.intrinsics(0.9)[None, None].repeat(b, v, 1, 1).clone()
    return dict(image=torch.rand(b, v, 3, hw, hw, generator=gen), extrinsics=extr, intrinsics=intr,
                near=torch.full((b, v), 1.0), far=torch.full((b, v), 30.0))


def encoder_goldens(out):
    """Goldens of the encoder-side modules from the reference's own PyTorch code (CPU).
    Stubbed third-party pieces (documented as UNPINNED in the modules): the DINO ViT (torch.hub, no network) is
    OUR restatement plugged into the reference's BackboneDino; e3nn's rotate_sh is OUR rotate_sh."""
    import helpers
    from types import SimpleNamespace
    import src.misc.sh_utils as ref_sh
    from latentsplat_b200.misc import sh_utils as our_sh
    from latentsplat_b200.model.encoder.backbone import dino_vit
    ref_sh.rotate_sh = our_sh.rotate_sh
    import src.model.encoder.common.gaussian_adapter as ref_ga
    ref_ga.rotate_sh = our_sh.rotate_sh
    torch.hub.load = lambda repo, model, *a, **k: dino_vit.build_dino(model)
    from src.global_cfg import set_cfg
    set_cfg(SimpleNamespace(dataset=SimpleNamespace(view_sampler=SimpleNamespace(num_context_views=2))))

    from src.geometry.epipolar_lines import get_depth, project_rays
    from src.model.encoder.backbone.backbone_dino import BackboneDinoCfg
    from src.model.encoder.common.gaussian_adapter import GaussianAdapterCfg
    from src.model.encoder.encoder_epipolar import EncoderEpipolar, EncoderEpipolarCfg, OpacityMappingCfg
    from src.model.encoder.epipolar.epipolar_transformer import EpipolarTransformerCfg
    from src.model.encoder.epipolar.image_self_attention import ImageSelfAttentionCfg
    from src.model.discriminator.discriminator_patch_gan import DiscriminatorPatchGan, DiscriminatorPatchGanCfg
    from fractions import Fraction

    # ---- geometry: project_rays / get_depth incl. rays that miss the other image ------------------------
    gen = torch.Generator().manual_seed(11)
    n = 400
    origins = torch.randn(n, 3, generator=gen) * 0.3
    directions = torch.nn.functional.normalize(torch.randn(n, 3, generator=gen) + torch.tensor([0.0, 0.0, 1.5]), dim=-1)
    extr = synthetic.pose(0.7, -8.0, 0.1, -0.05)[None].expand(n, 4, 4)
    intr = synthetic.intrinsics(0.8)[None].expand(n, 3, 3)
    pr = project_rays(origins, directions, extr, intr, torch.full((n,), 0.5), torch.full((n,), 20.0))
    pr2 = project_rays(origins, directions, extr, intr)
    xy = torch.rand(n, 2, generator=gen)
    dep = get_depth(origins, directions, xy, extr, intr)
    np.savez(out / "epipolar_geometry.npz", **{f"nf_{k}": v.numpy() for k, v in pr.items()},
             **{f"inf_{k}": v.numpy() for k, v in pr2.items()}, depth=dep.numpy())

    # ---- full encoder -----------------------------------------------------------------------------------
    sa, et, ga, enc = encoder_cfgs()
    cfg = EncoderEpipolarCfg(**enc, backbone=BackboneDinoCfg("dino", "dino_vitb8"), visualizer=None,
                             gaussian_adapter=GaussianAdapterCfg(**ga),
                             epipolar_transformer=EpipolarTransformerCfg(self_attention=ImageSelfAttentionCfg(**sa), **et),
                             opacity_mapping=OpacityMappingCfg(0.0, 0.0, 1))
    model = EncoderEpipolar(cfg, 3, 4, Fraction(1), True).eval()
    inventory = helpers.init_by_name(model, seed=3)
    ctx = encoder_context()
    with torch.no_grad():
        det = model(ctx, 0, deterministic=True)
        torch.manual_seed(123)
        sto = model(ctx, 0, deterministic=False)
    pack = lambda g, p: {f"{p}_means": g.means.numpy(), f"{p}_cov": g.covariances.numpy(), f"{p}_opac": g.opacities.numpy(),
                         f"{p}_csh": g.color_harmonics.numpy(), f"{p}_fsh": g.feature_harmonics.params.numpy()}
    sub = lambda d: {k: v[:, ::5] for k, v in d.items()}      # every 5th Gaussian keeps the fixture small
    np.savez_compressed(out / "encoder.npz", **sub(pack(det, "det")), **sub(pack(sto, "sto")),
                        det_sum=np.array([float(det.means.double().sum()), float(det.covariances.double().sum()),
                                          float(det.color_harmonics.double().sum()),
                                          float(det.feature_harmonics.params.double().sum())]),
                        inventory=np.array([f"{n}:{'x'.join(map(str, s))}" for n, s in inventory]))

    # ---- PatchGAN ----------------------------------------------------------------------------------------
    disc = DiscriminatorPatchGan(DiscriminatorPatchGanCfg("patch_gan", "kl_f8", pretrained=False), 3).eval()
    dinv = helpers.init_by_name(disc, seed=4)
    x = torch.rand(2, 3, 64, 64, generator=gen)
    with torch.no_grad():
        y = disc(x)
    np.savez(out / "patch_gan.npz", out=y.numpy(), inventory=np.array([f"{n}:{'x'.join(map(str, s))}" for n, s in dinv]))


def loss_inputs(seed: int = 23):
    """Seeded Prediction / GroundTruth fields + a toy "last layer": image = W (3x3 channel mix) applied to x, logits from the image."""
    gen = torch.Generator().manual_seed(seed)
    r = lambda *s: torch.rand(*s, generator=gen)
    return dict(W=torch.randn(3, 3, generator=gen) * 0.5 + torch.eye(3), x=r(2, 2, 3, 12, 16), gt_image=r(2, 2, 3, 12, 16),
                depth=r(2, 2, 12, 16) * 4.0 - 0.5, near=torch.tensor([[1.0, 0.8], [1.2, 1.0]]), far=torch.tensor([[20.0, 30.0], [15.0, 25.0]]),
                mean=torch.randn(2, 2, 4, 6, 8, generator=gen), logvar=torch.randn(2, 2, 4, 6, 8, generator=gen) * 3.0,
                logits_real=torch.randn(2, 2, 1, 3, 4, generator=gen), Wd=torch.randn(1, 3, generator=gen))


def loss_goldens(out):
    """tests/golden/losses.npz: every loss of /root/reference/src/loss and LossGroup's generator / discriminator passes incl. the
    adaptive weight, computed by the reference's own modules on seeded CPU inputs."""
    from src.loss import (LossDepthCfg, LossDiscriminatorCfg, LossGeneratorCfg, LossGroupCfg, LossKlCfg, LossL1Cfg, LossMseCfg,
                          get_loss_group)
    from src.model.diagonal_gaussian_distribution import DiagonalGaussianDistribution
    from src.model.types import GroundTruth, Prediction
    x = loss_inputs()
    W = x["W"].clone().requires_grad_(True)
    image = torch.einsum("oc,bvchw->bvohw", W, x["x"])
    logits_fake = torch.einsum("oc,bvchw->bvohw", x["Wd"], image)[..., ::4, ::4]
    pred = Prediction(image=image, posterior=DiagonalGaussianDistribution(x["mean"], x["logvar"], dim=2), depth=x["depth"],
                      logits_fake=logits_fake, logits_real=x["logits_real"])
    gt = GroundTruth(image=x["gt_image"], near=x["near"], far=x["far"])
    res = {}
    for tag, depth_cfg in (("plain", LossDepthCfg(weight=0.25)),
                           ("bilateral2", LossDepthCfg(weight=0.25, sigma_image=10.0, use_second_derivative=True))):
        for disc in ("hinge", "vanilla"):
            cfg = LossGroupCfg(nll=[LossMseCfg(weight=10.0), LossL1Cfg(weight=1.0), LossKlCfg(weight=1e-3), depth_cfg],
                               generator=LossGeneratorCfg(weight=0.5, apply_after_step=3),
                               discriminator=LossDiscriminatorCfg(weight=2.0, loss=disc, apply_after_step=3))
            group = get_loss_group("target", cfg)
            for step in (0, 7):
                total, d = group.forward_generator(pred, gt, step, last_layer_weights=W)
                key = f"{tag}_{disc}_s{step}"
                res[f"{key}_gen_total"] = total.detach().numpy()
                for k, v in d.items():
                    res[f"{key}_gen_{k}_u"] = v.unweighted.detach().numpy()
                    res[f"{key}_gen_{k}_w"] = v.weighted.detach().numpy()
                if step >= 3:
                    total, d = group.forward_discriminator(pred, gt, step)
                    res[f"{key}_dis_total"] = total.detach().numpy()
                    for k, v in d.items():
                        res[f"{key}_dis_{k}_u"] = v.unweighted.detach().numpy()
                        res[f"{key}_dis_{k}_w"] = v.weighted.detach().numpy()
    np.savez(out / "losses.npz", **res)
    print("wrote losses.npz with", len(res), "values")


def shim_inputs(seed: int = 31):
    gen = torch.Generator().manual_seed(seed)
    cams = torch.zeros(5, 18)
    cams[:, 0], cams[:, 1], cams[:, 2], cams[:, 3] = 0.9, 1.6, 0.5, 0.5
    for i in range(5):
        cams[i, 6:] = torch.linalg.inv(synthetic.pose(0.2 * i, 10.0 * i, 0.02 * i, -0.03))[:3].reshape(-1)
    return dict(images=torch.rand(2, 2, 3, 90, 160, generator=gen), intrinsics=synthetic.intrinsics(0.9)[None, None].repeat(2, 2, 1, 1).clone(),
                extrinsics=torch.stack([synthetic.pose(0.4, 25.0, 0.1, 0.2), synthetic.pose(-0.3, -10.0)])[None].repeat(2, 1, 1, 1), cameras=cams)


def shim_goldens(out):
    """tests/golden/dataset_shims.npz: crop shim (LANCZOS + centre crop), flip augmentation and RE10k pose conversion computed by the
    reference's own functions."""
    from src.dataset.shims.augmentation_shim import reflect_views
    from src.dataset.shims.crop_shim import rescale_and_crop
    x = shim_inputs()
    img, intr = rescale_and_crop(x["images"], x["intrinsics"], (64, 64))
    refl = reflect_views({"image": x["images"], "extrinsics": x["extrinsics"]})
    # convert_poses is a method that touches no state: call it unbound
    from src.dataset.dataset_re10k import DatasetRE10k
    ex, k = DatasetRE10k.convert_poses(None, x["cameras"])
    np.savez_compressed(out / "dataset_shims.npz", crop_image=img.numpy(), crop_intrinsics=intr.numpy(), flip_image=refl["image"].numpy(),
                        flip_extrinsics=refl["extrinsics"].numpy(), pose_extrinsics=ex.numpy(), pose_intrinsics=k.numpy())
    print("wrote dataset_shims.npz")


def main():
    install_reference()
    out = Path(__file__).resolve().parent
    if len(sys.argv) > 1 and sys.argv[1] == "losses":      # only the loss fixture (the others are unchanged)
        loss_goldens(out)
        return
    if len(sys.argv) > 1 and sys.argv[1] == "shims":
        shim_goldens(out)
        return

    # ---- SH ---------------------------------------------------------------------------
    from src.misc.sh_utils import eval_sh
    gen = torch.Generator().manual_seed(7)
    dirs = torch.randn(64, 3, generator=gen, dtype=torch.float64)
    dirs = dirs / dirs.norm(dim=-1, keepdim=True)
    sh = torch.randn(64, 3, 25, generator=gen, dtype=torch.float64)
    np.savez(out / "sh_eval.npz", dirs=dirs.numpy(), sh=sh.numpy(),
             **{f"deg{d}": eval_sh(d, sh, dirs).numpy() for d in range(5)})

    # ---- camera -----------------------------------------------------------------------
    from src.geometry.projection import get_fov
    from src.model.decoder.cuda_splatting import get_projection_matrix
    extr = torch.stack([synthetic.pose(0.3, 7.0, -0.1, 0.2), synthetic.pose(-1.0, -12.0, 0.4, -0.3), synthetic.pose()])
    intr = torch.stack([synthetic.intrinsics(0.86), synthetic.intrinsics(1.2), synthetic.intrinsics(0.7)])
    intr[1, 1, 1] = 1.5
    intr[2, 0, 2] = 0.45
    near, far = torch.tensor([1.0, 0.5, 2.0]), torch.tensor([100.0, 20.0, 7.0])
    fov = get_fov(intr)
    proj = get_projection_matrix(near, far, fov[:, 0], fov[:, 1])
    view_t = extr.inverse().transpose(1, 2)
    full = view_t @ proj.transpose(1, 2)
    np.savez(out / "camera.npz", extrinsics=extr.numpy(), intrinsics=intr.numpy(), near=near.numpy(),
             far=far.numpy(), fov=fov.numpy(), projection=proj.numpy(), view_t=view_t.numpy(), full_t=full.numpy())

    # ---- render_cuda / DecoderSplattingCUDA through the reference's own code -----------
    from src.model.decoder.decoder_splatting_cuda import DecoderSplattingCUDA, DecoderSplattingCUDACfg
    from src.model.types import Gaussians
    cfg = RENDER_CFG
    x = scene_inputs(cfg["seed"], cfg["G"], cfg["b"], cfg["v"])
    dec = DecoderSplattingCUDA(DecoderSplattingCUDACfg("splatting_cuda"), [0.2, 0.4, 0.6], variational=False)
    g = Gaussians(x["means"], x["covariances"], x["opacities"], x["color_harmonics"], x["feature_harmonics"])
    with torch.no_grad():
        o = dec.forward(g, x["extrinsics"], x["intrinsics"], x["near"], x["far"], (cfg["H"], cfg["W"]))
    np.savez_compressed(out / "render_cuda.npz", color=o.color.numpy(), feature_mean=o.feature_posterior.mean.numpy(),
                        feature_logvar=o.feature_posterior.logvar.numpy(), mask=o.mask.numpy(), depth=o.depth.numpy())
    encoder_goldens(out)
    loss_goldens(out)
    shim_goldens(out)
    print("wrote", [p.name for p in out.glob("*.npz")])


if __name__ == "__main__":
    main()
