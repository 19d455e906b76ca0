"""Depth <-> relative disparity (/root/reference/src/model/encoder/epipolar/conversions.py:5-27)."""
from torch import Tensor


def relative_disparity_to_depth(relative_disparity: Tensor, near: Tensor, far: Tensor, eps: float = 1e-10) -> Tensor:
    """0 -> near, 1 -> far."""
    disp_near, disp_far = 1 / (near + eps), 1 / (far + eps)
    return 1 / ((1 - relative_disparity) * (disp_near - disp_far) + disp_far + eps)


def depth_to_relative_disparity(depth: Tensor, near: Tensor, far: Tensor, eps: float = 1e-10) -> Tensor:
    disp_near, disp_far, disp = 1 / (near + eps), 1 / (far + eps), 1 / (depth + eps)
    return 1 - (disp - disp_far) / (disp_near - disp_far + eps)
