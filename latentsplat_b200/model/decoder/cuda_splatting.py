"""Host side of the splatting decoder: camera set-up and ONE batched rasterizer call.

Mirrors the public functions of /root/reference/src/model/decoder/cuda_splatting.py
(`get_projection_matrix` :19-46, `render_cuda` :56-167, `render_cuda_orthographic` :170-292,
`render_depth_cuda` :298-340, `RenderOutput` :49-54) with the same argument meaning, but:
  * the per-view Python loop (:124-162) is a single `rasterize_views` call over all views;
  * no `.item()` host syncs: tan(fov/2) stays on the device (:135-136);
  * the feature SH evaluation (:94-101) and the 1/near scene rescale (:75-82) are fused
    into the rasterizer's preprocess kernel (`feature_shs`, `scene_scale`);
  * `views_per_scene` lets one scene be rendered from several cameras without the
    materialised per-view copies of decoder_splatting_cuda.py:71-86.
"""
from __future__ import annotations

from dataclasses import dataclass
from math import isqrt
from typing import Literal, Optional

import torch
from torch import Tensor

from ...rasterizer import RasterDebug, rasterize_views
from latentsplat_b200.geometry.inverse import inv3x3, inv_affine4x4  # closed-form camera inverses (no cuSOLVER)


_FOV_RAYS: dict = {}


def _fov_rays(device) -> Tensor:
    """The four constant image-edge points of get_fov, created once per device (a per-call
    `torch.tensor(list, device=cuda)` is a pageable H2D copy, i.e. a stream sync on every step)."""
    key = str(device)
    if key not in _FOV_RAYS:
        _FOV_RAYS[key] = torch.tensor([[0.0, 0.5, 1.0], [1.0, 0.5, 1.0], [0.5, 0.0, 1.0], [0.5, 1.0, 1.0]],
                                      dtype=torch.float32, device=device)
    return _FOV_RAYS[key]


def get_fov(intrinsics: Tensor) -> Tensor:
    """Field of view (x, y) of normalised intrinsics; /root/reference/src/geometry/projection.py:233-247."""
    inv = inv3x3(intrinsics)
    rays = torch.einsum("bij,kj->kbi", inv, _fov_rays(intrinsics.device))   # left, right, top, bottom
    rays = rays / rays.norm(dim=-1, keepdim=True)
    return torch.stack(((rays[0] * rays[1]).sum(dim=-1).acos(), (rays[2] * rays[3]).sum(dim=-1).acos()), dim=-1)


def get_projection_matrix(near: Tensor, far: Tensor, fov_x: Tensor, fov_y: Tensor) -> Tensor:
    """x,y -> (-1,1), z -> (0,1), w = z_view (cuda_splatting.py:19-46)."""
    tan_x, tan_y = (0.5 * fov_x).tan(), (0.5 * fov_y).tan()
    top, right = tan_y * near, tan_x * near
    bottom, left = -top, -right
    out = torch.zeros((near.shape[0], 4, 4), dtype=torch.float32, device=near.device)
    out[:, 0, 0] = 2 * near / (right - left)
    out[:, 1, 1] = 2 * near / (top - bottom)
    out[:, 0, 2] = (right + left) / (right - left)
    out[:, 1, 2] = (top + bottom) / (top - bottom)
    out[:, 3, 2] = 1
    out[:, 2, 2] = far / (far - near)
    out[:, 2, 3] = -(far * near) / (far - near)
    return out


@dataclass
class RenderOutput:
    color: Optional[Tensor]    # (batch, 3, h, w)
    feature: Optional[Tensor]  # (batch, channels, h, w)
    mask: Tensor               # (batch, h, w)
    depth: Tensor              # (batch, h, w)


def _upper_triangle(cov: Tensor) -> Tensor:
    """(…,3,3) -> (…,6) in the order 00 01 02 11 12 22 (torch.triu_indices, cuda_splatting.py:148)."""
    return torch.stack((cov[..., 0, 0], cov[..., 0, 1], cov[..., 0, 2], cov[..., 1, 1], cov[..., 1, 2],
                        cov[..., 2, 2]), dim=-1)


def _camera_matrices(extrinsics: Tensor, near: Tensor, far: Tensor, fov_x: Tensor, fov_y: Tensor):
    """view (transposed), full projection (transposed) as at cuda_splatting.py:114-118."""
    projection = get_projection_matrix(near, far, fov_x, fov_y).transpose(1, 2)
    view = inv_affine4x4(extrinsics).transpose(1, 2)
    return view, view @ projection


def _split_sh(color_sh, feature_sh, use_sh):
    """Shared colour/feature argument handling of render_cuda (:84-107)."""
    kw = dict(sh_degree=0)
    if use_sh:
        if color_sh is not None:
            kw["sh_degree"] = isqrt(color_sh.shape[-1]) - 1
            kw["shs"] = color_sh.transpose(-1, -2)           # b g xyz n -> b g n xyz
        if feature_sh is not None:
            kw["feature_shs"] = feature_sh                    # evaluated in-kernel (0.5 + eval_sh)
    else:
        if color_sh is not None:
            kw["colors_precomp"] = color_sh[..., 0]
        if feature_sh is not None:
            kw["features"] = feature_sh[..., 0]
    return kw


def _raster_arguments(extrinsics, intrinsics, near, far, image_shape, background_color, gaussian_means,
                      gaussian_covariances, gaussian_opacities, gaussian_color_sh_coefficients,
                      gaussian_feature_sh_coefficients, scale_invariant, use_sh):
    """Camera set-up of render_cuda (:75-82, :108-118) -> positional/keyword arguments of rasterize_views."""
    assert gaussian_color_sh_coefficients is not None or gaussian_feature_sh_coefficients is not None
    assert use_sh or gaussian_color_sh_coefficients is None or gaussian_color_sh_coefficients.shape[-1] == 1
    scene_scale = None
    if scale_invariant:  # keep everything in a well-conditioned range (:75-82); means/cov scaled in-kernel
        scene_scale = 1 / near
        extrinsics = extrinsics.clone()
        extrinsics[..., :3, 3] = extrinsics[..., :3, 3] * scene_scale[:, None]
        near, far = near * scene_scale, far * scene_scale
    h, w = image_shape
    fov_x, fov_y = get_fov(intrinsics).unbind(dim=-1)
    tanfov = torch.stack(((0.5 * fov_x).tan(), (0.5 * fov_y).tan()), dim=-1)
    view, full_projection = _camera_matrices(extrinsics, near, far, fov_x, fov_y)
    args = (gaussian_means, _upper_triangle(gaussian_covariances), gaussian_opacities)
    kwargs = dict(viewmatrix=view, projmatrix=full_projection, campos=extrinsics[:, :3, 3], tanfov=tanfov,
                  image_height=h, image_width=w, bg=background_color, scene_scale=scene_scale,
                  **_split_sh(gaussian_color_sh_coefficients, gaussian_feature_sh_coefficients, use_sh))
    return args, kwargs


def render_cuda(extrinsics: Tensor, intrinsics: Tensor, near: Tensor, far: Tensor, image_shape: tuple[int, int],
                background_color: Tensor, gaussian_means: Tensor, gaussian_covariances: Tensor,
                gaussian_opacities: Tensor, gaussian_color_sh_coefficients: Optional[Tensor] = None,
                gaussian_feature_sh_coefficients: Optional[Tensor] = None, scale_invariant: bool = True,
                use_sh: bool = True, views_per_scene: int = 1, debug: Optional[RasterDebug] = None,
                capacity: Optional[int] = None) -> RenderOutput:
    """Same contract as the reference's `render_cuda`.  Camera tensors have `batch` = number of views V;
    Gaussian tensors have V // views_per_scene rows (views_per_scene=1 reproduces the reference call).
    `capacity`: see rasterize_views (None = exact, int = sync-free)."""
    assert extrinsics.shape[0] == gaussian_means.shape[0] * views_per_scene
    args, kwargs = _raster_arguments(extrinsics, intrinsics, near, far, image_shape, background_color,
                                     gaussian_means, gaussian_covariances, gaussian_opacities,
                                     gaussian_color_sh_coefficients, gaussian_feature_sh_coefficients,
                                     scale_invariant, use_sh)
    if PER_VIEW_LOOP:
        return _render_per_view(args, kwargs, views_per_scene)
    color, feature, alpha, depth, _ = rasterize_views(*args, debug=debug, capacity=capacity, **kwargs)
    return RenderOutput(color, feature, alpha, depth)


PER_VIEW_LOOP = False     # bench.py's GPU comparator: call the rasterizer once per view, like the reference's host loop

_PER_VIEW_KEYS = ("viewmatrix", "projmatrix", "campos", "tanfov", "bg", "scene_scale")


def _render_per_view(args, kwargs, views_per_scene: int) -> RenderOutput:
    """The reference's call structure (cuda_splatting.py:124-166): one rasterizer invocation per view with exact key-list
    sizing (one host sync each), outputs stacked afterwards.  Only used as the measured baseline of bench.py."""
    n_views = kwargs["viewmatrix"].shape[0]
    outs = []
    for i in range(n_views):
        s = i // views_per_scene
        a = tuple(t[s:s + 1] for t in args)
        kw = {}
        for k, v in kwargs.items():
            if isinstance(v, Tensor):
                kw[k] = v[i:i + 1] if k in _PER_VIEW_KEYS else v[s:s + 1]
            else:
                kw[k] = v
        outs.append(rasterize_views(*a, **kw))
    cat = lambda j: None if outs[0][j] is None else torch.cat([o[j] for o in outs])
    return RenderOutput(cat(0), cat(1), cat(2), cat(3))


def prepare_render_call(extrinsics, intrinsics, near, far, image_shape, background_color, gaussian_means,
                        gaussian_covariances, gaussian_opacities, gaussian_color_sh_coefficients=None,
                        gaussian_feature_sh_coefficients=None, scale_invariant=True, use_sh=True):
    """render_cuda's inputs -> an un-launched RasterCall (profiling: time the rasterizer stage by stage)."""
    from ...rasterizer import prepare_call
    args, kwargs = _raster_arguments(extrinsics, intrinsics, near, far, image_shape, background_color,
                                     gaussian_means, gaussian_covariances, gaussian_opacities,
                                     gaussian_color_sh_coefficients, gaussian_feature_sh_coefficients,
                                     scale_invariant, use_sh)
    return prepare_call(*args, **kwargs)


def render_cuda_orthographic(extrinsics: Tensor, width: Tensor, height: Tensor, near: Tensor, far: Tensor,
                             image_shape: tuple[int, int], background_features: Tensor, gaussian_means: Tensor,
                             gaussian_covariances: Tensor, gaussian_opacities: Tensor,
                             gaussian_color_sh_coefficients: Optional[Tensor] = None,
                             gaussian_feature_sh_coefficients: Optional[Tensor] = None, fov_degrees: float = 0.1,
                             use_sh: bool = True, dump: Optional[dict] = None) -> RenderOutput:
    """Fake orthographic view: a far-away camera with a tiny field of view (cuda_splatting.py:170-292)."""
    b = extrinsics.shape[0]
    h, w = image_shape
    fov_x = torch.tensor(fov_degrees, device=extrinsics.device).deg2rad()
    tan_x = (0.5 * fov_x).tan()
    distance_to_near = (0.5 * width) / tan_x
    tan_y = 0.5 * height / distance_to_near
    fov_y = (2 * tan_y).atan()
    near, far = near + distance_to_near, far + distance_to_near
    move_back = torch.eye(4, dtype=torch.float32, device=extrinsics.device).repeat(b, 1, 1)
    move_back[:, 2, 3] = -distance_to_near
    extrinsics = extrinsics @ move_back
    if dump is not None:
        dump.update(extrinsics=extrinsics, fov_x=fov_x, fov_y=fov_y, near=near, far=far)
    view, full_projection = _camera_matrices(extrinsics, near, far, fov_x.expand(b), fov_y)
    tanfov = torch.stack((tan_x.expand(b), tan_y.expand(b)), dim=-1)
    color, feature, alpha, depth, _ = rasterize_views(
        gaussian_means, _upper_triangle(gaussian_covariances), gaussian_opacities,
        viewmatrix=view, projmatrix=full_projection, campos=extrinsics[:, :3, 3], tanfov=tanfov,
        image_height=h, image_width=w, bg=background_features,
        **_split_sh(gaussian_color_sh_coefficients, gaussian_feature_sh_coefficients, use_sh))
    return RenderOutput(color, feature, alpha, depth)


DepthRenderingMode = Literal["depth", "disparity", "relative_disparity", "log"]


def depth_to_relative_disparity(depth: Tensor, near: Tensor, far: Tensor, eps: float = 1e-10) -> Tensor:
    """0 at near, 1 at far (/root/reference/src/model/encoder/epipolar/conversions.py:17-27)."""
    disp_near, disp_far, disp = 1 / (near + eps), 1 / (far + eps), 1 / (depth + eps)
    return 1 - (disp - disp_far) / (disp_near - disp_far + eps)


def render_depth_cuda(extrinsics: Tensor, intrinsics: Tensor, near: Tensor, far: Tensor, image_shape: tuple[int, int],
                      gaussian_means: Tensor, gaussian_covariances: Tensor, gaussian_opacities: Tensor,
                      scale_invariant: bool = True, mode: DepthRenderingMode = "depth",
                      views_per_scene: int = 1) -> Tensor:
    """Depth (or a transform of it) rendered as a fake colour (cuda_splatting.py:298-340)."""
    v = extrinsics.shape[0]
    means_v = gaussian_means.repeat_interleave(views_per_scene, dim=0) if views_per_scene > 1 else gaussian_means
    world2cam = inv_affine4x4(extrinsics)
    z = torch.einsum("bj,bgj->bg", world2cam[:, 2, :3], means_v) + world2cam[:, 2, 3:4]
    if mode == "disparity":
        z = 1 / z
    elif mode == "relative_disparity":
        z = depth_to_relative_disparity(z, near[:, None], far[:, None])
    elif mode == "log":
        z = z.minimum(near[:, None]).maximum(far[:, None]).log()
    covs = gaussian_covariances.repeat_interleave(views_per_scene, dim=0) if views_per_scene > 1 else gaussian_covariances
    opac = gaussian_opacities.repeat_interleave(views_per_scene, dim=0) if views_per_scene > 1 else gaussian_opacities
    result = render_cuda(extrinsics, intrinsics, near, far, image_shape,
                         torch.zeros((v, 3), dtype=z.dtype, device=z.device), means_v, covs, opac,
                         z[:, :, None, None].expand(-1, -1, 3, 1), scale_invariant=scale_invariant).color
    return result.mean(dim=1)
