"""GPU data shims used by `EncoderEpipolar.get_data_shim`
(/root/reference/src/dataset/shims/patch_shim.py:4-42, bounds_shim.py:9-80)."""
from __future__ import annotations

import torch
from torch import Tensor

from ...misc.cache import device_constant
from latentsplat_b200.geometry.inverse import inv2x2  # closed-form camera inverses (no cuSOLVER)


def apply_patch_shim_to_views(views: dict, patch_size: int) -> dict:
    """Centre-crop to a multiple of patch_size and rescale the normalised focal lengths."""
    _, _, _, h, w = views["image"].shape
    assert h % 2 == 0 and w % 2 == 0
    h_new, w_new = (h // patch_size) * patch_size, (w // patch_size) * patch_size
    row, col = (h - h_new) // 2, (w - w_new) // 2
    intrinsics = views["intrinsics"].clone()
    intrinsics[:, :, 0, 0] *= w / w_new
    intrinsics[:, :, 1, 1] *= h / h_new
    return {**views, "image": views["image"][:, :, :, row:row + h_new, col:col + w_new], "intrinsics": intrinsics}


def apply_patch_shim(batch: dict, patch_size: int) -> dict:
    return {**batch, "context": apply_patch_shim_to_views(batch["context"], patch_size),
            "target": apply_patch_shim_to_views(batch["target"], patch_size)}


def compute_depth_for_disparity(extrinsics: Tensor, intrinsics: Tensor, image_shape: tuple[int, int],
                                disparity: float, delta_min: float = 1e-6) -> Tensor:
    """Depth at which the largest camera baseline corresponds to `disparity` pixels."""
    origins = extrinsics[:, :, :3, 3]
    deltas = (origins[:, None, :, :] - origins[:, :, None, :]).norm(dim=-1).clip(min=delta_min)
    baselines = deltas.flatten(1).max(dim=1).values
    h, w = image_shape
    pixel_size = device_constant((1 / w, 1 / h), extrinsics.device)
    inv = inv2x2(intrinsics[..., :2, :2])
    pixel_size = torch.einsum("...ij,j->...i", inv, pixel_size)
    return baselines / (disparity * pixel_size.flatten(1).mean(dim=1))


def apply_bounds_shim(batch: dict, near_disparity: float, far_disparity: float) -> dict:
    context, target = batch["context"], batch["target"]
    _, cv, _, h, w = context["image"].shape
    near = compute_depth_for_disparity(context["extrinsics"], context["intrinsics"], (h, w), near_disparity)
    far = compute_depth_for_disparity(context["extrinsics"], context["intrinsics"], (h, w), far_disparity)
    tv = target["image"].shape[1]
    return {**batch,
            "context": {**context, "near": near[:, None].expand(-1, cv), "far": far[:, None].expand(-1, cv)},
            "target": {**target, "near": near[:, None].expand(-1, tv), "far": far[:, None].expand(-1, tv)}}
