"""Inference callers and formats around the render path.

  * `load_lightning_checkpoint`: a latentSplat Lightning `.ckpt` (ModelWrapper state dict: `encoder.*`, `decoder.*`,
    `autoencoder.*`, `discriminator.*`; /root/reference/src/main.py:138-146, model_wrapper.py:183-214) into a RenderPipeline --
    the parameter trees are name-identical (tests/test_encoder_cpu.py), so this is a strict load with the loss / LPIPS buffers of
    the wrapper dropped;
  * `EvaluationIndex`: the `assets/evaluation_index/*.json` files (scene -> list of {"context": [...], "target": [...]} entries,
    src/dataset/view_sampler/view_sampler_evaluation.py:18-60);
  * `predict_target_views`: `ModelWrapper.test_step` (model_wrapper.py:517-566): b = 1, stochastic Gaussians and latent sample,
    every target view of the entry, decoded by the VAE -- rendered in chunks of views so that a 300-frame video fits;
  * `save_predictions`: the PNG layout of test_step (`<root>/<scene>/<context indices>/color/<index:06d>.png`, image_io.py:37-70).
"""
from __future__ import annotations

import json
from dataclasses import dataclass
from fractions import Fraction
from pathlib import Path
from typing import Dict, Iterator, List, Mapping, Optional, Sequence, Tuple, Union

import torch
from torch import Tensor

from .pipeline import RenderPipeline, rescale

_PREFIXES = ("encoder", "decoder", "autoencoder", "discriminator")


def load_lightning_checkpoint(ckpt: Union[str, Path, Mapping[str, Tensor]], pipe: RenderPipeline, strict: bool = True) -> Dict[str, List[str]]:
    """Loads the four sub-modules of `pipe`; returns {"missing": [...], "unexpected": [...], "ignored": [...]}.
    `ignored` = wrapper-only entries (loss groups' buffers, step counters); strict=True raises on missing / unexpected keys."""
    state = ckpt
    if not isinstance(ckpt, Mapping):
        state = torch.load(ckpt, map_location="cpu", weights_only=True)
    state = state.get("state_dict", state)
    parts: Dict[str, Dict[str, Tensor]] = {p: {} for p in _PREFIXES}
    ignored = []
    for key, value in state.items():
        head, _, rest = key.partition(".")
        if head in parts and rest:
            parts[head][rest] = value
        else:
            ignored.append(key)
    missing, unexpected = [], []
    for name in _PREFIXES:
        module = getattr(pipe, name, None)
        if module is None:
            if parts[name]:
                ignored.extend(f"{name}.{k}" for k in parts[name])
            continue
        if not parts[name] and not list(module.state_dict()):
            continue
        res = module.load_state_dict(parts[name], strict=False)
        missing += [f"{name}.{k}" for k in res.missing_keys]
        unexpected += [f"{name}.{k}" for k in res.unexpected_keys]
    if strict and (missing or unexpected):
        raise RuntimeError(f"checkpoint does not match the pipeline: missing {missing[:5]}{'...' if len(missing) > 5 else ''}, "
                           f"unexpected {unexpected[:5]}{'...' if len(unexpected) > 5 else ''}")
    return {"missing": missing, "unexpected": unexpected, "ignored": ignored}


@dataclass
class IndexEntry:
    context: Tuple[int, ...]
    target: Tuple[int, ...]


class EvaluationIndex:
    def __init__(self, path: Union[str, Path]):
        raw = json.loads(Path(path).read_text())
        # a scene maps to a list of entries (or null when the scene has no valid view pair; older files hold one dict)
        self.index: Dict[str, List[IndexEntry]] = {}
        for scene, entries in raw.items():
            if entries is None:
                continue
            if isinstance(entries, dict):
                entries = [entries]
            self.index[scene] = [IndexEntry(tuple(e["context"]), tuple(e["target"])) for e in entries]

    def __len__(self) -> int:
        return sum(len(v) for v in self.index.values())

    def __iter__(self) -> Iterator[Tuple[str, IndexEntry]]:
        for scene, entries in self.index.items():
            for e in entries:
                yield scene, e

    def entries(self, scene: str) -> List[IndexEntry]:
        return self.index.get(scene, [])


@torch.no_grad()
def predict_target_views(pipe: RenderPipeline, batch: dict, views_per_chunk: Optional[int] = None, deterministic: bool = False) -> Tensor:
    """test_step for one example (b = 1): encoder once, then the target views in chunks -> (v, 3, H, W) decoded images.
    `deterministic=True` is the validation / video "deterministic" branch (mode of the Gaussians and of the latent posterior)."""
    context, target = batch["context"], batch["target"]
    if context["image"].shape[0] != 1:
        raise ValueError("predict_target_views renders one example at a time (test_step asserts b == 1)")
    size = tuple(context["image"].shape[-2:]) if "image_shape" not in batch else batch["image_shape"]
    gaussians = pipe.encoder(context, 0, features=None, deterministic=deterministic)
    if pipe.variational not in ("gaussians", "none"):
        g = gaussians.flatten()
    else:
        g = gaussians.mode() if deterministic else gaussians.sample()
    v = target["extrinsics"].shape[1]
    step = v if not views_per_chunk else max(1, int(views_per_chunk))
    images = []
    for lo in range(0, v, step):
        sl = slice(lo, min(v, lo + step))
        out = pipe.decoder(g, target["extrinsics"][:, sl], target["intrinsics"][:, sl], target["near"][:, sl], target["far"][:, sl], size)
        latent = out.feature_posterior.mode() if deterministic else out.feature_posterior.sample()
        z = rescale(latent, Fraction(1, pipe.supersampling_factor))
        skip_z = None
        if pipe.autoencoder.expects_skip:
            skip_z = torch.cat((out.color.detach(), latent), dim=-3) if pipe.autoencoder.expects_skip_extra else latent
        images.append(pipe.autoencoder.decode(z, skip_z)[0])
    return torch.cat(images, dim=0)


def to_uint8(image: Tensor):
    """(3 | 1, H, W) float in [0, 1] -> (H, W, 3) uint8 numpy (image_io.py:37-56: clip, * 255, truncate)."""
    if image.dim() == 2:
        image = image[None]
    if image.shape[0] == 1:
        image = image.expand(3, -1, -1)
    return (image.detach().clip(0, 1) * 255).to(torch.uint8).permute(1, 2, 0).cpu().numpy()


def save_predictions(images: Tensor, indices: Sequence[int], root: Union[str, Path], scene: str, context_indices: Sequence[int]) -> List[Path]:
    from PIL import Image
    folder = Path(root) / scene / "_".join(str(i) for i in sorted(int(i) for i in context_indices)) / "color"
    folder.mkdir(parents=True, exist_ok=True)
    paths = []
    for index, image in zip(indices, images):
        path = folder / f"{int(index):0>6}.png"
        Image.fromarray(to_uint8(image)).save(path)
        paths.append(path)
    return paths
