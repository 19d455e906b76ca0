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
    from helpers_oracle_stub import OracleGaussianRasterizer, OracleSettings
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


def main():
    install_reference()
    out = Path(__file__).resolve().parent

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
    print("wrote", [p.name for p in out.glob("*.npz")])


if __name__ == "__main__":
    main()
