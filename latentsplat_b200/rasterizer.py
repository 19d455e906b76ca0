"""Differentiable Gaussian rasterizer on libls_raster.so (sm_100a), plus the
`diff_gaussian_rasterization` API the reference imports at
/root/reference/src/model/decoder/cuda_splatting.py:6-9 and calls at :132-158.

Two entry points:
  * `rasterize_views(...)`  -- batched: V views of S scenes in one launch sequence
    (what our decoder uses; replaces the Python loop at cuda_splatting.py:124-162);
  * `GaussianRasterizationSettings` / `GaussianRasterizer` -- the per-view drop-in.

No CPU path exists: tensors must live on a CUDA device and the library must be built.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import NamedTuple, Optional

import torch
from torch import Tensor, nn

from . import _capi

__all__ = ["GaussianRasterizationSettings", "GaussianRasterizer", "rasterize_views", "RasterDebug", "RasterCall",
           "RasterCapacityError", "check_overflow", "reserve_overflow_slots", "reset_overflow_guard"]


class RasterCapacityError(RuntimeError):
    """A sync-free (`capacity=`) rasterizer call needed more (tile, Gaussian) slots than it was given: the affected tiles were
    rendered from truncated lists, so the images / gradients of that call are wrong."""


class _OverflowGuard:
    """Makes the overflow flag of the sync-free mode (stats[2], ls_raster.h) an error instead of a number somebody may read.

    Every capacity-mode forward enqueues a 16-byte copy of its `stats` into pinned host memory right behind its kernels (no
    synchronisation).  The copies are examined whenever the host next touches the rasterizer -- the next forward or backward
    call, `GraphedStep.replay()`, or an explicit `check_overflow()` -- and the first completed one with the flag set raises
    RasterCapacityError: at most one step late, never silently.  Inside a CUDA-graph capture the copy becomes a node of the
    graph writing to a buffer that lives as long as the process, so every replay refreshes it."""

    def __init__(self):
        self.pending = []       # eager: (event, pinned stats, capacity)
        self.resident = []      # captured in a CUDA graph: (pinned stats, capacity)
        self.free = []

    def reserve(self, n: int) -> None:
        """Pinned buffers cannot be allocated during a stream capture: GraphedStep reserves some beforehand."""
        while len(self.free) < n:
            self.free.append(torch.zeros(4, dtype=torch.int32).pin_memory())

    def watch(self, stats: Tensor, capacity: int) -> None:
        if torch.cuda.is_current_stream_capturing():
            if not self.free:
                raise RuntimeError("sync-free rasterizer call inside a CUDA-graph capture without reserved overflow slots: "
                                   "call latentsplat_b200.rasterizer.reserve_overflow_slots() before capturing")
            host = self.free.pop()
            host.copy_(stats, non_blocking=True)
            self.resident.append((host, capacity))
            return
        host = self.free.pop() if self.free else torch.zeros(4, dtype=torch.int32).pin_memory()
        host.copy_(stats, non_blocking=True)
        event = torch.cuda.Event()
        event.record(torch.cuda.current_stream(stats.device))
        self.pending.append((event, host, capacity))

    @staticmethod
    def _raise(host: Tensor, capacity: int) -> None:
        needed = int(host[0])
        host[2] = 0
        raise RasterCapacityError(f"rasterizer key capacity overflow: a call needed {needed} (tile, Gaussian) slots but was given "
                                  f"{capacity}; its images and gradients are truncated -- raise `capacity` / recalibrate")

    def poll(self, block: bool = False) -> None:
        if torch.cuda.is_current_stream_capturing():         # no event queries during a capture; the replay path polls
            return
        keep = []
        failed = None
        for event, host, capacity in self.pending:
            if block:
                event.synchronize()
            if not event.query():
                keep.append((event, host, capacity))
            elif int(host[2]) and failed is None:
                failed = (host, capacity)
            else:
                self.free.append(host)
        self.pending = keep
        if failed is None:
            for host, capacity in self.resident:
                if int(host[2]):
                    failed = (host, capacity)
                    break
        if failed is not None:
            self._raise(*failed)


_GUARD = _OverflowGuard()


def reserve_overflow_slots(n: int = 8) -> None:
    """Pre-allocate the pinned flag buffers that sync-free calls captured into a CUDA graph will write to."""
    _GUARD.reserve(n)


def reset_overflow_guard() -> None:
    """Forget every outstanding / graph-resident overflow flag (after the caller has dealt with an overflow, e.g. re-captured its
    CUDA graph with a larger capacity: the old graph's flag buffer would otherwise keep reporting)."""
    _GUARD.free.extend(h for _, h, _ in _GUARD.pending)
    _GUARD.pending.clear()
    _GUARD.resident.clear()


def check_overflow(block: bool = False) -> None:
    """Raise RasterCapacityError if any completed sync-free rasterizer call overflowed its key capacity.  `block=True` first waits
    for the outstanding calls (use at the end of an epoch / before trusting a result); the default never synchronises."""
    _GUARD.poll(block)


def _ptr(t: Optional[Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def _f32c(t: Optional[Tensor], name: str) -> Optional[Tensor]:
    if t is None:
        return None
    if not t.is_cuda:
        raise RuntimeError(f"{name} must be a CUDA tensor: latentsplat_b200 has no CPU fallback")
    return t.detach().to(torch.float32).contiguous()


class RasterDebug:
    """Optional sink: pass `debug=RasterDebug()` to keep the binning state for parity dumps."""
    def __init__(self):
        self.state = None
        self.num_rendered = None
        self.stats = None   # device int32 (4): num_rendered, longest tile list, overflow flag, 0


class _Buffers:
    """All device buffers of one forward (owned by torch, handed to the C ABI as pointers)."""

    def __init__(self, scene: _capi.LsRasterScene, device, sort_smem_keys: int = 0):
        lib = _capi.load()
        sz = _capi.LsRasterSizes()
        _capi.check(lib.ls_raster_sizes(C.byref(scene), C.byref(sz)), "ls_raster_sizes")
        self.sizes = sz
        V, G, H, W = scene.n_views, scene.G, scene.H, scene.W
        e = lambda *shape, dtype=torch.float32: torch.empty(shape, dtype=dtype, device=device)
        self.geom = e(V, G, 8)
        self.chan = e(V, G, sz.chan_stride)
        self.radii = e(V, G, dtype=torch.int32)
        self.tiles_touched = e(V, G, dtype=torch.int32)
        self.clamped = e(V, G, dtype=torch.uint8)
        self.tile_count = e(sz.tile_slots, dtype=torch.int32)
        self.tile_offsets = e(sz.tile_slots + 1, dtype=torch.int32)
        self.stats = torch.zeros(4, dtype=torch.int32, device=device)
        self.keys = None
        self.keys_tmp = None
        self.sorted_cull = None     # (capacity, 4): x, y, extents, id   -- the tile queues in list order, written by the sort
        self.sorted_rec = None      # (capacity, rec_stride): geometry + channel record
        self.capacity = 0
        self.final_T = e(V, H, W)
        self.n_contrib = e(V, H, W, dtype=torch.int32)
        self.sort_smem_keys = sort_smem_keys

    def alloc_keys(self, capacity: int, device):
        self.capacity = int(capacity)
        n = max(self.capacity, 1)
        self.keys = torch.empty(n, dtype=torch.int64, device=device)
        self.keys_tmp = torch.empty(n, dtype=torch.int64, device=device)
        self.sorted_cull = torch.empty((n, 4), dtype=torch.float32, device=device)
        self.sorted_rec = torch.empty((n, self.sizes.rec_stride), dtype=torch.float32, device=device)

    def as_struct(self) -> _capi.LsRasterState:
        return _capi.LsRasterState(
            _ptr(self.geom), _ptr(self.chan), _ptr(self.radii), _ptr(self.tiles_touched), _ptr(self.clamped),
            _ptr(self.tile_count), _ptr(self.tile_offsets), _ptr(self.stats), _ptr(self.keys), _ptr(self.keys_tmp),
            self.capacity, _ptr(self.final_T), _ptr(self.n_contrib), self.sizes.chan_stride, self.sort_smem_keys,
            _ptr(self.sorted_cull), _ptr(self.sorted_rec), self.sizes.rec_stride, 0)


def _make_scene(V, vps, G, H, W, Cf, color_mode, sh_degree, feature_mode, fdeg, means3D, cov3D, opacity, color,
                feature, viewmatrix, projmatrix, campos, tanfov, bg, scene_scale) -> _capi.LsRasterScene:
    return _capi.LsRasterScene(V, vps, G, H, W, Cf, color_mode, sh_degree, feature_mode, fdeg, _ptr(means3D),
                               _ptr(cov3D), _ptr(opacity), _ptr(color), _ptr(feature), _ptr(viewmatrix),
                               _ptr(projmatrix), _ptr(campos), _ptr(tanfov), _ptr(bg), _ptr(scene_scale))


class RasterCall:
    """One batched rasterizer invocation: owns the prepared (contiguous fp32) inputs, the state buffers and
    the output images, and launches the C-ABI stages.  `_Rasterize` drives it for autograd; bench.py /
    profiling drive it stage by stage to bracket individual kernels with CUDA events."""

    def __init__(self, means3D, cov3D, opacity, color, feature, viewmatrix, projmatrix, campos, tanfov, bg,
                 scene_scale, H, W, views_per_scene, color_mode, sh_degree, feature_mode, feature_sh_degree,
                 sort_smem_keys=0, capacity=None):
        self.lib = _capi.load()
        self.capacity = capacity
        self.device = means3D.device
        self.means3D = _f32c(means3D, "means3D")
        self.cov3D = _f32c(cov3D, "cov3D")
        self.opacity = _f32c(opacity, "opacities")
        self.color = _f32c(color, "colour")
        self.feature = _f32c(feature, "features")
        self.viewmatrix, self.projmatrix = _f32c(viewmatrix, "viewmatrix"), _f32c(projmatrix, "projmatrix")
        self.campos, self.tanfov, self.bg = _f32c(campos, "campos"), _f32c(tanfov, "tanfov"), _f32c(bg, "bg")
        self.scene_scale = _f32c(scene_scale, "scene_scale")
        S, G = self.means3D.shape[0], self.means3D.shape[1]
        V = self.viewmatrix.shape[0]
        if V != S * views_per_scene:
            raise RuntimeError(f"{V} views != {S} scenes x {views_per_scene} views per scene")
        Cf = 0 if self.feature is None else self.feature.shape[2]
        self.cfg = (H, W, views_per_scene, color_mode, sh_degree, feature_mode, feature_sh_degree, Cf, V, S, G)
        self.scene = _make_scene(V, views_per_scene, G, H, W, Cf, color_mode, sh_degree, feature_mode,
                                 feature_sh_degree, self.means3D, self.cov3D, self.opacity, self.color,
                                 self.feature, self.viewmatrix, self.projmatrix, self.campos, self.tanfov, self.bg,
                                 self.scene_scale)
        dev = self.device
        with torch.cuda.device(dev):
            self.buf = _Buffers(self.scene, dev, sort_smem_keys)
            self.images = dict(
                color=torch.empty((V, 3, H, W), device=dev) if color_mode != _capi.COLOR_NONE else None,
                feature=torch.empty((V, Cf, H, W), device=dev) if Cf else None,
                alpha=torch.empty((V, H, W), device=dev), depth=torch.empty((V, H, W), device=dev))
        self.num_rendered = None

    def _stream(self):
        return torch.cuda.current_stream(self.device).cuda_stream

    def forward_stage(self, stages: int) -> None:
        im = _capi.LsRasterImages(_ptr(self.images["color"]), _ptr(self.images["feature"]),
                                  _ptr(self.images["alpha"]), _ptr(self.images["depth"]))
        st = self.buf.as_struct()
        with torch.cuda.device(self.device):
            _capi.check(self.lib.ls_raster_forward(C.byref(self.scene), C.byref(st), C.byref(im), stages,
                                                   self._stream()), f"ls_raster_forward(stages={stages})")
        # preprocess + scan | scatter | sort | blend
        _capi.KERNEL_LAUNCHES[0] += 2 * bool(stages & 1) + bool(stages & 2) + bool(stages & 4) + bool(stages & 8)

    def size_keys(self) -> int:
        """The one host sync of a step: exact size of the key lists (the lineage syncs once per VIEW for
        the same number; here once per batched call)."""
        self.num_rendered = int(self.buf.stats[0].item())
        self.buf.alloc_keys(self.num_rendered, self.device)
        return self.num_rendered

    def forward(self):
        _GUARD.poll()                       # earlier sync-free calls that overflowed raise here (never silently truncated)
        if self.capacity is None:           # exact: one 4-byte D2H sync per batched call
            self.forward_stage(_capi.STAGE_GEOMETRY)
            self.size_keys()
            self.forward_stage(_capi.STAGE_RENDER)
        else:                               # sync-free (CUDA-graph capturable): caller-chosen capacity;
            self.buf.alloc_keys(int(self.capacity), self.device)   # stats[2] flags an overflow -> _OverflowGuard
            self.forward_stage(_capi.STAGE_ALL)
            with torch.cuda.device(self.device):
                _GUARD.watch(self.buf.stats, int(self.capacity))
        return self.images

    @staticmethod
    def _padded_rows(like: Tensor, dense: bool = False):
        """Gradient buffer for per-Gaussian coefficient rows (S, G, a, b): rows of a*b floats padded to a multiple of 8 floats
        (whole 32-byte sectors, see LsRasterGrads.color_grad_pitch).  Returns (buffer (S, G, pitch), view shaped like `like`,
        pitch); rows of <= 4 floats (precomputed colours / features) stay dense."""
        S, G = like.shape[:2]
        row = like[0, 0].numel()
        if dense or like.dim() != 4 or row <= 4 or row % 8 == 0:
            buf = torch.empty_like(like)
            return buf, buf, 0
        pitch = (row + 7) // 8 * 8
        buf = torch.empty((S, G, pitch), dtype=like.dtype, device=like.device)
        return buf, buf[:, :, :row].unflatten(2, tuple(like.shape[2:])), pitch

    def alloc_grads(self, want_means2D: bool):
        H, W, vps, color_mode, sh_degree, feature_mode, fdeg, Cf, V, S, G = self.cfg
        dev = self.device
        with torch.cuda.device(dev):
            cbuf = cview = fbuf = fview = None
            self.color_pitch = self.feature_pitch = 0
            # the specialised coefficient backward moves a warp's rows as one bulk copy and wants them dense
            dense = bool(self.lib.ls_raster_dense_sh_grads(C.byref(self.scene)))
            if self.color is not None:
                cbuf, cview, self.color_pitch = self._padded_rows(self.color, dense)
            if self.feature is not None:
                fbuf, fview, self.feature_pitch = self._padded_rows(self.feature, dense)
            self.grad_buf = dict(color=cbuf, feature=fbuf)
            self.grad_out = dict(
                record=torch.empty((V, G, self.buf.sizes.grad_stride), device=dev),
                means3D=torch.empty_like(self.means3D), cov3D=torch.empty_like(self.cov3D),
                opacity=torch.empty_like(self.opacity), color=cview, feature=fview,
                means2D=torch.empty((V, G, 3), device=dev) if want_means2D else None)
        return self.grad_out

    def backward_stage(self, stages: int, g_color, g_feature, g_alpha, g_depth) -> None:
        o = self.grad_out
        grads = _capi.LsRasterGrads(_ptr(g_color), _ptr(g_feature), _ptr(g_alpha), _ptr(g_depth), _ptr(o["record"]),
                                    self.buf.sizes.grad_stride, 0, _ptr(o["means3D"]), _ptr(o["cov3D"]),
                                    _ptr(o["opacity"]), _ptr(self.grad_buf["color"]), _ptr(self.grad_buf["feature"]),
                                    _ptr(o["means2D"]), self.color_pitch, self.feature_pitch)
        st = self.buf.as_struct()
        with torch.cuda.device(self.device):
            _capi.check(self.lib.ls_raster_backward(C.byref(self.scene), C.byref(st), C.byref(grads), stages,
                                                    self._stream()), f"ls_raster_backward(stages={stages})")
        _capi.KERNEL_LAUNCHES[0] += bool(stages & 1) + bool(stages & 2)


class _Rasterize(torch.autograd.Function):
    @staticmethod
    def forward(ctx, means3D, cov3D, opacity, color, feature, means2D, viewmatrix, projmatrix, campos, tanfov, bg,
                scene_scale, H, W, views_per_scene, color_mode, sh_degree, feature_mode, feature_sh_degree,
                sort_smem_keys, capacity, debug):
        call = RasterCall(means3D, cov3D, opacity, color, feature, viewmatrix, projmatrix, campos, tanfov, bg,
                          scene_scale, H, W, views_per_scene, color_mode, sh_degree, feature_mode,
                          feature_sh_degree, sort_smem_keys, capacity)
        images = call.forward()
        ctx.call = call
        ctx.has_means2D = means2D is not None
        if debug is not None:
            debug.state = call.buf
            debug.num_rendered = call.num_rendered
            debug.stats = call.buf.stats
        radii = call.buf.radii
        ctx.mark_non_differentiable(radii)
        return images["color"], images["feature"], images["alpha"], images["depth"], radii

    @staticmethod
    def backward(ctx, g_color, g_feature, g_alpha, g_depth, _g_radii):
        call: RasterCall = ctx.call
        _GUARD.poll()
        gc = lambda t: None if t is None else t.to(torch.float32).contiguous()
        o = call.alloc_grads(ctx.has_means2D)
        call.backward_stage(_capi.BWD_ALL, gc(g_color), gc(g_feature), gc(g_alpha), gc(g_depth))
        return (o["means3D"], o["cov3D"], o["opacity"], o["color"], o["feature"], o["means2D"]) + (None,) * 16


def _color_sh_mode() -> int:
    """Which real-SH convention the colour `shs` use.  The reference passes them to its fork's CUDA (cuda_splatting.py:91,146),
    whose source is absent [EXT]; the two candidates are the in-tree src/misc/sh_utils.py basis (default, what its features use)
    and the graphdeco 3DGS coefficient order.  LS_SH_BASIS=intree|3dgs picks one; read per call so tests can flip it."""
    basis = os.environ.get("LS_SH_BASIS", "intree").lower()
    if basis not in ("intree", "3dgs"):
        raise ValueError(f"LS_SH_BASIS must be 'intree' or '3dgs', got {basis!r}")
    return _capi.COLOR_SH_3DGS if basis == "3dgs" else _capi.COLOR_SH


def _normalize(means3D, cov3D, opacities, *, viewmatrix, projmatrix, campos, tanfov, image_height, image_width,
               bg=None, shs=None, colors_precomp=None, features=None, feature_shs=None, sh_degree=0,
               scene_scale=None, sort_smem_keys=0, capacity=None):
    """Argument checking shared by rasterize_views and prepare_call; returns RasterCall's positional args."""
    if shs is not None and colors_precomp is not None:
        raise ValueError("Please provide only one of either SHs or precomputed colors!")
    if features is not None and feature_shs is not None:
        raise ValueError("Provide either pre-evaluated features or feature SH coefficients, not both")
    V = viewmatrix.shape[0]
    S = means3D.shape[0]
    if S == 0 or V % S:
        raise ValueError(f"{V} views cannot be split over {S} scenes")
    if shs is not None:
        color, color_mode = shs, _color_sh_mode()
        n = shs.shape[2]
        if (sh_degree + 1) ** 2 > n:
            raise ValueError(f"sh_degree {sh_degree} needs {(sh_degree + 1) ** 2} coefficients, got {n}")
        if n != (sh_degree + 1) ** 2:
            color = shs[:, :, : (sh_degree + 1) ** 2]
    elif colors_precomp is not None:
        color, color_mode = colors_precomp, _capi.COLOR_PRECOMP
    else:
        color, color_mode = None, _capi.COLOR_NONE
    fdeg = 0
    if feature_shs is not None:
        feature, feature_mode = feature_shs, _capi.FEATURE_SH
        nf = feature_shs.shape[3]
        fdeg = int(round(nf ** 0.5)) - 1
        if (fdeg + 1) ** 2 != nf:
            raise ValueError(f"feature SH coefficient count {nf} is not a square")
    elif features is not None:
        feature, feature_mode = features, _capi.FEATURE_PRECOMP
    else:
        feature, feature_mode = None, _capi.FEATURE_NONE
    if color is None and feature is None:
        raise ValueError("nothing to render: provide colours and/or features (cuda_splatting.py:71)")
    if bg is None:
        bg = torch.zeros((V, 3), device=means3D.device)
    return (means3D, cov3D, opacities, color, feature, viewmatrix.reshape(V, 16), projmatrix.reshape(V, 16), campos,
            tanfov, bg, scene_scale, int(image_height), int(image_width), V // S, color_mode, int(sh_degree),
            feature_mode, fdeg, int(sort_smem_keys), None if capacity is None else int(capacity))


def prepare_call(*args, **kwargs) -> RasterCall:
    """Same arguments as rasterize_views (minus means2D/debug); returns an un-launched RasterCall so that a
    profiler can run and time the stages one by one (no autograd)."""
    return RasterCall(*_normalize(*args, **kwargs))


def rasterize_views(means3D: Tensor, cov3D: Tensor, opacities: Tensor, *, viewmatrix: Tensor, projmatrix: Tensor,
                    campos: Tensor, tanfov: Tensor, image_height: int, image_width: int,
                    bg: Optional[Tensor] = None, shs: Optional[Tensor] = None,
                    colors_precomp: Optional[Tensor] = None, features: Optional[Tensor] = None,
                    feature_shs: Optional[Tensor] = None, sh_degree: int = 0, scene_scale: Optional[Tensor] = None,
                    means2D: Optional[Tensor] = None, sort_smem_keys: int = 0, capacity: Optional[int] = None,
                    debug: Optional[RasterDebug] = None):
    """Render V = S * views_per_scene views in one launch sequence.

    means3D (S,G,3), cov3D (S,G,6), opacities (S,G); colour: `shs` (S,G,n,3) or
    `colors_precomp` (S,G,3) or neither; features: `features` (S,G,C) pre-evaluated or
    `feature_shs` (S,G,C,nf) evaluated in-kernel as 0.5+eval_sh (cuda_splatting.py:94-101).
    Cameras (V,...): viewmatrix/projmatrix (V,4,4) transposed as in cuda_splatting.py:115-118,
    campos (V,3), tanfov (V,2), bg (V,3), scene_scale (V,) or None.
    `capacity`: None = exact sizing of the key lists (one 4-byte host sync per call); an int = sync-free
    mode with that many (tile, Gaussian) slots -- `debug.stats[2]` is 1 if the scene needed more (the excess
    entries of the affected tiles are dropped), `debug.stats[0]` is the number it needed.
    Returns (color|None (V,3,H,W), feature|None (V,C,H,W), alpha (V,H,W), depth (V,H,W), radii (V,G)).

    More than LS_MAX_VALUE_CHANNELS (16) blended channels -- colour(3) + C, e.g. the variational kl_f16 / kl_f32 latents
    (2*16+3 = 35, 2*64+3 = 131; the reference ships those autoencoder configs) -- are rendered in several passes of <= 16
    channels over the same Gaussians: pass 0 carries the colour, alpha and depth, later passes only feature channels (their
    alpha / depth outputs are dropped, so those gradients reach the geometry once).  Every pass repeats preprocess + sort (cost:
    one rasterizer call per 16 channels); gradients of the shared inputs add up through autograd.
    """
    n_color = 0 if (shs is None and colors_precomp is None) else 3
    feats = feature_shs if feature_shs is not None else features
    if feats is not None and n_color + feats.shape[2] > _capi.MAX_VALUE_CHANNELS:
        common = dict(viewmatrix=viewmatrix, projmatrix=projmatrix, campos=campos, tanfov=tanfov, image_height=image_height,
                      image_width=image_width, bg=bg, sh_degree=sh_degree, scene_scale=scene_scale,
                      sort_smem_keys=sort_smem_keys, capacity=capacity)
        key = "feature_shs" if feature_shs is not None else "features"
        first = _capi.MAX_VALUE_CHANNELS - n_color
        bounds = [0, first] + list(range(first + _capi.MAX_VALUE_CHANNELS, feats.shape[2], _capi.MAX_VALUE_CHANNELS)) + [feats.shape[2]]
        bounds = sorted(set(bounds))
        color = alpha = depth = radii = None
        parts = []
        for i, (lo, hi) in enumerate(zip(bounds[:-1], bounds[1:])):
            out = rasterize_views(means3D, cov3D, opacities, shs=shs if i == 0 else None,
                                  colors_precomp=colors_precomp if i == 0 else None, means2D=means2D if i == 0 else None,
                                  debug=debug if i == 0 else None, **{key: feats[:, :, lo:hi]}, **common)
            if i == 0:
                color, _, alpha, depth, radii = out
            parts.append(out[1])
        return color, torch.cat(parts, dim=1), alpha, depth, radii
    a = _normalize(means3D, cov3D, opacities, viewmatrix=viewmatrix, projmatrix=projmatrix, campos=campos,
                   tanfov=tanfov, image_height=image_height, image_width=image_width, bg=bg, shs=shs,
                   colors_precomp=colors_precomp, features=features, feature_shs=feature_shs, sh_degree=sh_degree,
                   scene_scale=scene_scale, sort_smem_keys=sort_smem_keys, capacity=capacity)
    return _Rasterize.apply(*a[:5], means2D, *a[5:], debug)


# ------------------------------------------------------------------------------------------
# diff_gaussian_rasterization drop-in (per-view API)
# ------------------------------------------------------------------------------------------
class GaussianRasterizationSettings(NamedTuple):
    """The 12 fields the reference passes at cuda_splatting.py:132-145."""
    image_height: int
    image_width: int
    tanfovx: float
    tanfovy: float
    bg: Tensor
    scale_modifier: float
    viewmatrix: Tensor
    projmatrix: Tensor
    sh_degree: int
    campos: Tensor
    prefiltered: bool
    debug: bool


def _cov3d_from_scales_rotations(scales: Tensor, rotations: Tensor, scale_modifier: float) -> Tensor:
    """[EXT] computeCov3D: Sigma = R S S^T R^T with quaternion (r,x,y,z); returns the 6 upper-triangle
    entries.  Plain torch: the reference never takes this path (it always passes cov3D_precomp)."""
    q = rotations / rotations.norm(dim=-1, keepdim=True)
    r, x, y, z = q.unbind(-1)
    R = torch.stack([1 - 2 * (y * y + z * z), 2 * (x * y - r * z), 2 * (x * z + r * y),
                     2 * (x * y + r * z), 1 - 2 * (x * x + z * z), 2 * (y * z - r * x),
                     2 * (x * z - r * y), 2 * (y * z + r * x), 1 - 2 * (x * x + y * y)], dim=-1).reshape(-1, 3, 3)
    M = R * (scales * scale_modifier)[:, None, :]
    cov = M @ M.transpose(1, 2)
    i, j = torch.triu_indices(3, 3)
    return cov[:, i, j]


class GaussianRasterizer(nn.Module):
    """Per-view rasterizer with the call signature used at cuda_splatting.py:150-158.

    Returns the 5-tuple (image | None, feature_map | None, mask (1,H,W), depth_map (1,H,W), radii (G,)).
    `shs`, `colors_precomp` and `features` may each be None; unlike stock 3DGS, passing neither colour input
    is legal as long as features are given (model_wrapper.py:369 `return_colors=False`).
    """

    def __init__(self, raster_settings: GaussianRasterizationSettings):
        super().__init__()
        self.raster_settings = raster_settings

    def forward(self, means3D, means2D, opacities, shs=None, colors_precomp=None, features=None, scales=None,
                rotations=None, cov3D_precomp=None):
        rs = self.raster_settings
        if shs is not None and colors_precomp is not None:
            raise Exception("Please provide only one of either SHs or precomputed colors!")
        if (scales is None or rotations is None) == (cov3D_precomp is None):
            raise Exception("Please provide exactly one of either scale/rotation pair or precomputed 3D covariance!")
        if not means3D.is_cuda:
            raise RuntimeError("GaussianRasterizer needs CUDA tensors: latentsplat_b200 has no CPU fallback")
        if cov3D_precomp is None:
            cov3D_precomp = _cov3d_from_scales_rotations(scales, rotations, rs.scale_modifier)
        dev = means3D.device
        if isinstance(rs.tanfovx, Tensor) or isinstance(rs.tanfovy, Tensor):  # cuda_splatting.py:260-261
            tanfov = torch.stack([torch.as_tensor(rs.tanfovx, device=dev).float().reshape(()),
                                  torch.as_tensor(rs.tanfovy, device=dev).float().reshape(())]).reshape(1, 2)
        else:
            tanfov = torch.tensor([[float(rs.tanfovx), float(rs.tanfovy)]], dtype=torch.float32, device=dev)
        color, feat, alpha, depth, radii = rasterize_views(
            means3D[None], cov3D_precomp[None], opacities.reshape(1, -1),
            viewmatrix=rs.viewmatrix.reshape(1, 4, 4), projmatrix=rs.projmatrix.reshape(1, 4, 4),
            campos=rs.campos.reshape(1, 3), tanfov=tanfov, image_height=rs.image_height,
            image_width=rs.image_width, bg=rs.bg.reshape(1, 3),
            shs=None if shs is None else shs[None], colors_precomp=None if colors_precomp is None else colors_precomp[None],
            features=None if features is None else features[None], sh_degree=rs.sh_degree,
            means2D=None if means2D is None else means2D[None])
        return (None if color is None else color[0], None if feat is None else feat[0], alpha, depth, radii[0])
