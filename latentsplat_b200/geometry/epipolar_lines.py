"""Epipolar segment clipping with the semantics of /root/reference/src/geometry/epipolar_lines.py.

`project_rays` clips each world ray to [near, far] and to the frame of another camera and returns the
image-space segment.  The reference selects between the four (near valid?, far valid?) cases with boolean
masked assignments (epipolar_lines.py:239-249) -- one host sync each; here the same selection is a pair of
`torch.where`s, so the whole function is a static launch sequence.
"""
from __future__ import annotations

from typing import Optional

import torch
from torch import Tensor

from .projection import get_world_rays, homogenize_points, homogenize_vectors, intersect_rays, matvec, project_camera_space
from latentsplat_b200.geometry.inverse import inv_affine4x4  # closed-form camera inverses (no cuSOLVER)


def _is_in_bounds(xy: Tensor, epsilon: float = 1e-6) -> Tensor:
    return (xy >= -epsilon).all(dim=-1) & (xy <= 1 + epsilon).all(dim=-1)


def _is_in_front_of_camera(xyz: Tensor, epsilon: float = 1e-6) -> Tensor:
    return xyz[..., -1] > -epsilon


def _is_positive_t(t: Tensor, epsilon: float = 1e-6) -> Tensor:
    return t > -epsilon


def _intersect_image_coordinate(intrinsics: Tensor, origins: Tensor, directions: Tensor, dim: int,
                                coordinate_value: float) -> dict:
    """Intersection of a camera-space ray's projection with the image-frame line x|y = value
    (epipolar_lines.py:56-110)."""
    other = 1 - dim
    fs, fo = intrinsics[..., dim, dim], intrinsics[..., other, other]
    cs, co = intrinsics[..., dim, 2], intrinsics[..., other, 2]
    os_, oo, oz = origins[..., dim], origins[..., other], origins[..., 2]
    ds, do, dz = directions[..., dim], directions[..., other], directions[..., 2]
    c = (coordinate_value - cs) / fs
    t = (c * oz - os_) / (ds - c * dz)                      # infinite t is fine
    coordinate_other = co + fo * (oo * (c * dz - ds) + do * (os_ - c * oz)) / (dz * os_ - ds * oz)
    coordinate_same = torch.ones_like(coordinate_other) * coordinate_value
    xy = torch.stack((coordinate_same, coordinate_other) if dim == 0 else (coordinate_other, coordinate_same), dim=-1)
    xyz = origins + t[..., None] * directions
    return {"t": t, "xy": xy, "valid": _is_in_bounds(xy) & _is_in_front_of_camera(xyz) & _is_positive_t(t)}


def _compare_projections(intersections, reduction: str) -> dict:
    """Pick, per ray, the valid intersection with the smallest / largest t (epipolar_lines.py:113-139)."""
    t = torch.stack([i["t"] for i in intersections])
    xy = torch.stack([i["xy"] for i in intersections])
    valid = torch.stack([i["valid"] for i in intersections])
    lowest = torch.inf if reduction == "min" else -torch.inf
    t = torch.where(valid, t, torch.full_like(t, lowest))
    reduced, selector = getattr(t, reduction)(dim=0)
    return {"t": reduced,
            "xy": xy.gather(0, selector[None, ..., None].expand(1, *selector.shape, 2))[0],
            "valid": valid.gather(0, selector[None])[0]}


def _compute_point_projection(xyz: Tensor, t: Tensor, intrinsics: Tensor) -> dict:
    xy = project_camera_space(xyz, intrinsics)
    return {"t": t, "xy": xy, "valid": _is_in_bounds(xy) & _is_in_front_of_camera(xyz) & _is_positive_t(t)}


def project_rays(origins: Tensor, directions: Tensor, extrinsics: Tensor, intrinsics: Tensor,
                 near: Optional[Tensor] = None, far: Optional[Tensor] = None, epsilon: float = 1e-6) -> dict:
    """World rays -> the image-space segment they trace in the camera (extrinsics, intrinsics)
    (epipolar_lines.py:157-251).  Returns t_min, t_max, xy_min, xy_max, overlaps_image."""
    world_to_cam = inv_affine4x4(extrinsics)
    origins = matvec(world_to_cam, homogenize_points(origins))[..., :3]
    directions = matvec(world_to_cam, homogenize_vectors(directions))[..., :3]

    frame = (_intersect_image_coordinate(intrinsics, origins, directions, 0, 0.0),
             _intersect_image_coordinate(intrinsics, origins, directions, 0, 1.0),
             _intersect_image_coordinate(intrinsics, origins, directions, 1, 0.0),
             _intersect_image_coordinate(intrinsics, origins, directions, 1, 1.0))
    frame_min = _compare_projections(frame, "min")
    frame_max = _compare_projections(frame, "max")

    if near is None:
        # projection at zero depth; an origin at the camera uses its direction instead (:194-212)
        mask_depth_zero = origins[..., -1] < epsilon
        mask_at_camera = origins.norm(dim=-1) < epsilon
        origins_for_projection = torch.where(mask_at_camera[..., None], directions, origins)
        at_zero = _compute_point_projection(origins_for_projection, torch.zeros_like(frame_min["t"]), intrinsics)
        at_zero["valid"] = at_zero["valid"] & ~(mask_depth_zero & ~mask_at_camera)
    else:
        at_zero = _compute_point_projection(origins + near[..., None] * directions,
                                            near.broadcast_to(frame_min["t"].shape), intrinsics)
    if far is None:
        at_inf = _compute_point_projection(directions, torch.ones_like(frame_min["t"]) * torch.inf, intrinsics)
    else:
        at_inf = _compute_point_projection(origins + far[..., None] * directions,
                                           far.broadcast_to(frame_min["t"].shape), intrinsics)

    # near end: the clipped point if it projects inside the frame, else the first frame crossing; same for far
    min_ok, max_ok = at_zero["valid"], at_inf["valid"]
    pick = lambda ok, a, b, k: torch.where(ok[..., None] if a[k].dim() > ok.dim() else ok, a[k], b[k])
    lo = {k: pick(min_ok, at_zero, frame_min, k) for k in ("t", "xy", "valid")}
    hi = {k: pick(max_ok, at_inf, frame_max, k) for k in ("t", "xy", "valid")}
    return {"t_min": lo["t"], "t_max": hi["t"], "xy_min": lo["xy"], "xy_max": hi["xy"],
            "overlaps_image": lo["valid"] & hi["valid"]}


def lift_to_3d(origins: Tensor, directions: Tensor, xy: Tensor, extrinsics: Tensor, intrinsics: Tensor) -> Tensor:
    """3D point on the ray (origins, directions) seen at image position xy of camera (extrinsics, intrinsics)
    (epipolar_lines.py:264-277)."""
    xy_origins, xy_directions = get_world_rays(xy, extrinsics, intrinsics)
    return intersect_rays(origins, directions, xy_origins, xy_directions)


def get_depth(origins: Tensor, directions: Tensor, xy: Tensor, extrinsics: Tensor, intrinsics: Tensor) -> Tensor:
    """Distance along the ray of the point seen at xy in the other camera (epipolar_lines.py:280-292)."""
    return (lift_to_3d(origins, directions, xy, extrinsics, intrinsics) - origins).norm(dim=-1)
