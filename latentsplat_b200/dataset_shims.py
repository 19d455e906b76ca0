"""Input side of the render path: the per-example shims and the RE10k chunk decoding that feed
`EncoderEpipolar` -- host code (they run in the data-loader workers), restated from
/root/reference/src/dataset/shims/crop_shim.py:11-93, augmentation_shim.py:8-37 and src/dataset/dataset_re10k.py:147-206.

  decode_images      JPEG bytes (uint8 tensors of an RE10k `.torch` chunk) -> (n, 3, h, w) float in [0, 1]
  convert_poses      the 18-float camera rows (fx fy cx cy 0 0 | 3x4 world-to-camera) -> camera-to-world 4x4 + normalised K
  rescale_and_crop   LANCZOS down-scale (through uint8, as the reference) to cover the target shape, centre crop, focal fix-up
  reflect_views      the horizontal-flip augmentation (image flip + mirrored extrinsics)
  make_example       one training / evaluation example from a chunk entry: baseline normalisation, bounds, augmentation, crop
"""
from __future__ import annotations

from io import BytesIO
from typing import Optional, Sequence, Tuple

import numpy as np
import torch
from torch import Tensor

from .geometry.inverse import inv_affine4x4


def decode_images(images: Sequence[Tensor]) -> Tensor:
    from PIL import Image
    out = []
    for raw in images:
        img = Image.open(BytesIO(raw.numpy().tobytes()))
        arr = np.asarray(img.convert("RGB") if img.mode != "RGB" else img, dtype=np.uint8)
        out.append(torch.from_numpy(arr.copy()).permute(2, 0, 1).to(torch.float32) / 255.0)       # torchvision ToTensor semantics
    return torch.stack(out)


def convert_poses(poses: Tensor) -> Tuple[Tensor, Tensor]:
    b = poses.shape[0]
    intrinsics = torch.eye(3, dtype=torch.float32).repeat(b, 1, 1)
    intrinsics[:, 0, 0], intrinsics[:, 1, 1] = poses[:, 0], poses[:, 1]
    intrinsics[:, 0, 2], intrinsics[:, 1, 2] = poses[:, 2], poses[:, 3]
    w2c = torch.eye(4, dtype=torch.float32).repeat(b, 1, 1)
    w2c[:, :3] = poses[:, 6:].reshape(b, 3, 4)
    return inv_affine4x4(w2c), intrinsics


def _lanczos(image: Tensor, shape: Tuple[int, int]) -> Tensor:
    from PIL import Image
    h, w = shape
    u8 = (image * 255).clip(0, 255).to(torch.uint8).permute(1, 2, 0).cpu().numpy()
    out = np.asarray(Image.fromarray(u8).resize((w, h), Image.LANCZOS)) / 255
    return torch.tensor(out, dtype=image.dtype, device=image.device).permute(2, 0, 1)


def center_crop(images: Tensor, intrinsics: Tensor, shape: Tuple[int, int]) -> Tuple[Tensor, Tensor]:
    h_in, w_in = images.shape[-2:]
    h_out, w_out = shape
    row, col = (h_in - h_out) // 2, (w_in - w_out) // 2
    intrinsics = intrinsics.clone()
    intrinsics[..., 0, 0] *= w_in / w_out
    intrinsics[..., 1, 1] *= h_in / h_out
    return images[..., row:row + h_out, col:col + w_out], intrinsics


def rescale_and_crop(images: Tensor, intrinsics: Tensor, shape: Tuple[int, int]) -> Tuple[Tensor, Tensor]:
    *batch, c, h_in, w_in = images.shape
    h_out, w_out = shape
    assert h_out <= h_in and w_out <= w_in
    s = max(h_out / h_in, w_out / w_in)
    hs, ws = round(h_in * s), round(w_in * s)
    assert hs == h_out or ws == w_out
    scaled = torch.stack([_lanczos(img, (hs, ws)) for img in images.reshape(-1, c, h_in, w_in)]).reshape(*batch, c, hs, ws)
    return center_crop(scaled, intrinsics, shape)


def apply_crop_shim(example: dict, shape: Tuple[int, int]) -> dict:
    def crop(views):
        image, intr = rescale_and_crop(views["image"], views["intrinsics"], shape)
        return {**views, "image": image, "intrinsics": intr}
    return {**example, "context": crop(example["context"]), "target": crop(example["target"])}


def reflect_extrinsics(extrinsics: Tensor) -> Tensor:
    m = torch.eye(4, dtype=torch.float32, device=extrinsics.device)
    m[0, 0] = -1
    return m @ extrinsics @ m


def reflect_views(views: dict) -> dict:
    return {**views, "image": views["image"].flip(-1), "extrinsics": reflect_extrinsics(views["extrinsics"])}


def apply_augmentation_shim(example: dict, generator: Optional[torch.Generator] = None) -> dict:
    if torch.rand((), generator=generator) < 0.5:              # keep the example as it is with probability 1/2
        return example
    return {**example, "context": reflect_views(example["context"]), "target": reflect_views(example["target"])}


def make_example(entry: dict, context_indices: Tensor, target_indices: Tensor, image_shape: Tuple[int, int], near: float, far: float,
                 make_baseline_1: bool = True, baseline_epsilon: float = 1e-3, augment: bool = False,
                 generator: Optional[torch.Generator] = None) -> Optional[dict]:
    """One example from an RE10k chunk entry ({"key", "cameras" (n, 18), "images" [uint8 JPEG tensors]}); None when the entry is
    skipped the way the reference skips it (wrong image size, insufficient baseline)."""
    extrinsics, intrinsics = convert_poses(entry["cameras"])
    ctx_img = decode_images([entry["images"][int(i)] for i in context_indices])
    tgt_img = decode_images([entry["images"][int(i)] for i in target_indices])
    if ctx_img.shape[1:] != (3, 360, 640) or tgt_img.shape[1:] != (3, 360, 640):
        return None
    scale = 1.0
    if len(context_indices) == 2 and make_baseline_1:
        a, b = extrinsics[context_indices][:, :3, 3]
        scale = (a - b).norm()
        if scale < baseline_epsilon:
            return None
        extrinsics = extrinsics.clone()
        extrinsics[:, :3, 3] /= scale
    def views(idx, img):
        n = len(idx)
        return {"extrinsics": extrinsics[idx], "intrinsics": intrinsics[idx], "image": img,
                "near": torch.full((n,), near, dtype=torch.float32) / scale, "far": torch.full((n,), far, dtype=torch.float32) / scale, "index": idx}
    sample = {"context": views(context_indices, ctx_img), "target": views(target_indices, tgt_img), "scene": entry["key"]}
    if augment:
        sample = apply_augmentation_shim(sample, generator)
    return apply_crop_shim(sample, image_shape)
