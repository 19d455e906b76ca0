"""Epipolar-sampling encoder: context images + cameras -> variational Gaussians.

Same constructor, parameter tree, forward signature and outputs as
/root/reference/src/model/encoder/encoder_epipolar.py:51-268.  Two data-flow changes, both exact:
  * `backbone_projection` (ReLU + Linear d_backbone -> d_feature, :70-73, applied per pixel at :143-145) runs on
    the un-repeated token grid and the result is repeated, instead of repeating 512 channels to full resolution
    first (pointwise ops commute with nearest-neighbour replication; the global token is added before the
    projection on the coarse grid, exactly as the reference adds it before its repeat is consumed);
  * all geometry below is sync-free (see latentsplat_b200/geometry).
"""
from __future__ import annotations

from dataclasses import dataclass, field
from fractions import Fraction
from typing import Literal, Optional

import torch
from torch import Tensor, nn

from ...geometry.projection import sample_image_grid
from ...misc.cache import device_constant
from ..diagonal_gaussian_distribution import DiagonalGaussianDistribution
from ..types import VariationalGaussians
from .backbone import Backbone, BackboneCfg, get_backbone
from .backbone.backbone_dino import BackboneDino
from .common.gaussian_adapter import GaussianAdapter, GaussianAdapterCfg
from .encoder import Encoder
from .epipolar.depth_predictor_monocular import DepthPredictorMonocular
from .epipolar.epipolar_transformer import EpipolarTransformer, EpipolarTransformerCfg
from .shims import apply_bounds_shim, apply_patch_shim
from latentsplat_b200.gemm import Linear, grouped_linear  # nn.Linear / F.linear with tcgen05 TF32 GEMMs on CUDA
from latentsplat_b200 import gaussian_head  # fused depth sampling + Gaussian adapter (sm_100a kernel)
from latentsplat_b200.conv import Conv2d  # tcgen05 implicit-GEMM convolutions (NHWC) with fused bias + activation


@dataclass
class OpacityMappingCfg:
    initial: float
    final: float
    warm_up: int


@dataclass
class EncoderEpipolarCfg:
    name: Literal["epipolar"]
    d_backbone: int
    d_feature: int
    num_monocular_samples: int
    num_surfaces: int
    predict_opacity: bool
    backbone: BackboneCfg
    near_disparity: float
    gaussian_adapter: GaussianAdapterCfg
    apply_bounds_shim: bool
    epipolar_transformer: EpipolarTransformerCfg
    opacity_mapping: OpacityMappingCfg
    gaussians_per_pixel: int
    use_epipolar_transformer: bool
    use_transmittance: bool
    visualizer: Optional[dict] = None        # EncoderVisualizerEpipolarCfg in the reference (out of scope)
    num_context_views: int = 2               # get_cfg().dataset.view_sampler.num_context_views in the reference


FOLD_HARMONICS = True     # fold SH masking + camera-to-world SH rotation into the to_gaussians weights (set False for A/B)


class EncoderEpipolar(Encoder[EncoderEpipolarCfg]):
    def __init__(self, cfg: EncoderEpipolarCfg, d_in: int, n_feature_channels: int, scale_factor: Fraction,
                 variational: bool) -> None:
        super().__init__(cfg, variational)
        self.backbone: Backbone = get_backbone(cfg.backbone, d_in, cfg.d_backbone, scale_factor)
        self.backbone_projection = nn.Sequential(nn.ReLU(), Linear(cfg.d_backbone, cfg.d_feature))
        self.epipolar_transformer = EpipolarTransformer(cfg.epipolar_transformer, cfg.d_feature, cfg.num_context_views) \
            if cfg.use_epipolar_transformer else None
        self.depth_predictor = DepthPredictorMonocular(cfg.d_feature, cfg.num_monocular_samples, cfg.num_surfaces,
                                                       cfg.use_transmittance)
        # NOTE twice the feature channels when variational: mean and log-variance (:87-90)
        self.gaussian_adapter = GaussianAdapter(cfg.gaussian_adapter,
                                                2 * n_feature_channels if variational else n_feature_channels)
        if cfg.predict_opacity:
            self.to_opacity = nn.Sequential(nn.ReLU(), Linear(cfg.d_feature, 1), nn.Sigmoid())
        self.to_gaussians = nn.Sequential(nn.ReLU(), Linear(cfg.d_feature,
                                                               cfg.num_surfaces * (2 + self.gaussian_adapter.d_in)))
        # the high-resolution skip only exists without downscaling (:104-111)
        self.high_resolution_skip = nn.Sequential(Conv2d(3, cfg.d_feature, 7, 1, 3, act="relu"), nn.Identity()) \
            if scale_factor == 1 else None

    def map_pdf_to_opacity(self, pdf: Tensor, global_step: int) -> Tensor:
        """https://www.desmos.com/calculator/opvwti3ba9 (:113-126)"""
        cfg = self.cfg.opacity_mapping
        x = cfg.initial + min(global_step / cfg.warm_up, 1) * (cfg.final - cfg.initial)
        exponent = 2 ** x
        return 0.5 * (1 - (1 - pdf) ** exponent + pdf ** (1 / exponent))

    def _backbone_features(self, images: Tensor) -> tuple[Tensor, int]:
        """(bv, c, H, W) -> (projected features (bv, d_feature, h, w), repeats): backbone_projection(backbone(x)) == the
        returned map with every pixel replicated repeats x repeats times (repeats = 1: already full resolution)."""
        if isinstance(self.backbone, BackboneDino) and self.backbone.cfg.upscale_mode == "repeat":
            local, glob = self.backbone.forward_tokens(images)
            x = self.backbone_projection((local + glob).permute(0, 2, 3, 1)).permute(0, 3, 1, 2)
            return x, self.backbone.n_repeats
        x = self.backbone(images)
        return self.backbone_projection(x.permute(0, 2, 3, 1)).permute(0, 3, 1, 2), 1

    def _gaussian_head(self, features: Tensor, c2w_rotations: Tensor, color_major: bool = False) -> Tensor:
        """to_gaussians(features) (:97-103) -- with the SH masking and the camera-to-world SH rotation of the Gaussian
        adapter (gaussian_adapter.py:44-61, 104-105) folded into the Linear: both are linear maps T_view of the raw
        output, so  T_v (W x + b) = (T_v W) x + T_v b  and each context view gets its own effective weight."""
        if not FOLD_HARMONICS:
            return self.to_gaussians(features)
        act, lin = self.to_gaussians[0], self.to_gaussians[1]
        b, v = features.shape[:2]
        srf, d = self.cfg.num_surfaces, 2 + self.gaussian_adapter.d_in
        t = torch.zeros((b, v, srf * d, srf * d), dtype=features.dtype, device=features.device)
        t_adapter = self.gaussian_adapter.harmonics_transform(c2w_rotations)                    # (b, v, d_in, d_in)
        for k in range(srf):
            o = k * d
            t[:, :, o, o] = 1.0
            t[:, :, o + 1, o + 1] = 1.0
            t[:, :, o + 2:o + d, o + 2:o + d] = t_adapter
        if color_major:
            # emit the colour SH block coefficient-major ((n, xyz) instead of (xyz, n)): that is the layout the rasterizer reads
            # (cuda_splatting.py:91), so the fused tail can hand out a transposed VIEW and no (G, 3, 25) <-> (G, 25, 3) copy is made
            # in either direction.  A row permutation of T, i.e. of the effective weight -- free.
            n = self.gaussian_adapter.d_color_sh
            perm = torch.arange(srf * d, device=t.device)
            for k in range(srf):
                o = k * d + 9                                                                   # 2 offsets + 3 scales + 4 quaternion
                perm[o:o + 3 * n] = o + torch.arange(3 * n, device=t.device).view(3, n).t().reshape(-1)
            t = t[:, :, perm]
        weight = t @ lin.weight                                                                 # (b, v, srf*d, d_feature)
        bias = (t @ lin.bias[:, None])[..., 0]
        return grouped_linear(act(features).flatten(0, 1), weight.flatten(0, 1), bias.flatten(0, 1)).unflatten(0, (b, v))

    def forward(self, context: dict, global_step: int, features: Optional[Tensor] = None,
                deterministic: bool = False, visualization_dump: Optional[dict] = None) -> VariationalGaussians:
        b, v = context["image"].shape[:2]
        images = context["image"].flatten(0, 1)
        features, repeats = self._backbone_features(images if features is None else features)
        device = features.device
        h, w = features.shape[-2] * repeats, features.shape[-1] * repeats
        features = features.unflatten(0, (b, v))

        sampling = None
        if self.epipolar_transformer is not None:
            # the un-replicated token grid goes in: the transformer's down-scaler collapses the replication algebraically
            features, sampling = self.epipolar_transformer(features, context["extrinsics"], context["intrinsics"],
                                                           context["near"], context["far"], repeats=repeats)
        elif repeats > 1:
            features = features.repeat_interleave(repeats, dim=3).repeat_interleave(repeats, dim=4)
        if self.high_resolution_skip is not None:
            features = features + self.high_resolution_skip(images).unflatten(0, (b, v))

        features = features.flatten(3).transpose(2, 3)                     # b v c h w -> b v (h w) c
        fused = self._fused_tail(features, context, global_step, deterministic, (h, w), visualization_dump)
        if fused is not None:
            return fused
        depths, densities = self.depth_predictor(features, context["near"], context["far"], deterministic,
                                                 1 if deterministic else self.cfg.gaussians_per_pixel)

        xy_ray, _ = sample_image_grid((h, w), device)
        xy_ray = xy_ray.reshape(h * w, 1, 2)
        gaussians = self._gaussian_head(features, context["extrinsics"][..., :3, :3])
        gaussians = gaussians.unflatten(-1, (self.cfg.num_surfaces, -1))                       # ... (srf c) -> ... srf c
        offset_xy = gaussians[..., :2].sigmoid()
        pixel_size = device_constant((1 / w, 1 / h), device)
        xy_ray = xy_ray + (offset_xy - 0.5) * pixel_size
        gpp = self.cfg.gaussians_per_pixel
        g = self.gaussian_adapter(context["extrinsics"][:, :, None, None, None], context["intrinsics"][:, :, None, None, None],
                                  xy_ray[..., None, :], depths, self.map_pdf_to_opacity(densities, global_step) / gpp,
                                  gaussians[..., None, 2:], (h, w), harmonics_ready=FOLD_HARMONICS)

        if visualization_dump is not None:
            visualization_dump["depth"] = depths.unflatten(2, (h, w))
            visualization_dump["scales"] = g.scales.flatten(1, 4)
            visualization_dump["rotations"] = g.rotations.flatten(1, 4)
            if sampling is not None:
                visualization_dump["sampling"] = sampling

        opacity_multiplier = self.to_opacity(features)[..., None] if self.cfg.predict_opacity else 1   # b v r () ()
        feature_harmonics = g.feature_harmonics.flatten(1, 4)              # b v r srf spp c d -> b (v r srf spp) c d
        feature_harmonics = DiagonalGaussianDistribution(
            **{"params" if self.variational else "mean": feature_harmonics}, dim=-2)
        return VariationalGaussians(g.means.flatten(1, 4), g.covariances.flatten(1, 4),
                                    (opacity_multiplier * g.opacities).flatten(1, 4),
                                    g.color_harmonics.flatten(1, 4), feature_harmonics)

    def _fused_tail(self, features: Tensor, context: dict, global_step: int, deterministic: bool, image_shape,
                    visualization_dump: Optional[dict]) -> Optional[VariationalGaussians]:
        """depth predictor -> xy offsets -> opacity mapping -> Gaussian adapter as ONE kernel each way
        (latentsplat_b200.gaussian_head) after the two Linear heads; None when the configuration needs the explicit sequence
        below (CPU tensors, visualisation dumps, hooks on the pdf / offset modules, transmittance or opacity heads)."""
        dp, ga, cfg = self.depth_predictor, self.gaussian_adapter, self.cfg
        spp = 1 if deterministic else cfg.gaussians_per_pixel
        if (visualization_dump is not None or not FOLD_HARMONICS or dp.to_pdf._forward_hooks or dp.to_offset._forward_hooks
                or not (gaussian_head.ENABLED and features.is_cuda and features.dtype == torch.float32)):
            return None
        b, v, r, _ = features.shape
        dlog = dp.projection(features)                                                          # (b, v, r, 2 * buckets)
        if not gaussian_head.supported(dlog, dlog, dp.num_samples, cfg.num_surfaces, spp, dp.use_transmittance, cfg.predict_opacity):
            return None
        raw = self._gaussian_head(features, context["extrinsics"][..., :3, :3], color_major=True)   # (b, v, r, 2 + adapter.d_in)
        # the reference's draw: torch.rand over (*pdf.shape[:-1], num_samples) = (b, v, r, srf, spp)
        u = None if deterministic else torch.rand((b, v, r, cfg.num_surfaces, spp), device=features.device).reshape(b, v, r, spp)
        om = cfg.opacity_mapping
        exponent = 2.0 ** (om.initial + min(global_step / om.warm_up, 1) * (om.final - om.initial))
        d_color, d_feature = 3 * ga.d_color_sh, ga.n_feature_channels * ga.d_feature_sh
        means, cov, opac, csh, fsh, _ = gaussian_head.gaussian_head(
            dlog, raw, u, context["extrinsics"], context["intrinsics"], context["near"], context["far"], image_shape, spp,
            d_color, d_feature, ga.cfg.gaussian_scale_min, ga.cfg.gaussian_scale_max, exponent, cfg.gaussians_per_pixel)
        feature_harmonics = fsh.unflatten(-1, (ga.n_feature_channels, ga.d_feature_sh))
        feature_harmonics = DiagonalGaussianDistribution(
            **{"params" if self.variational else "mean": feature_harmonics}, dim=-2)
        # csh rows are coefficient-major (see _gaussian_head): (b, G, 3, n) as the API wants it is a transposed view
        return VariationalGaussians(means, cov, opac, csh.unflatten(-1, (ga.d_color_sh, 3)).transpose(-1, -2), feature_harmonics)

    def get_data_shim(self):
        def data_shim(batch: dict) -> dict:
            et = self.cfg.epipolar_transformer
            batch = apply_patch_shim(batch, patch_size=et.self_attention.patch_size * et.downscale)
            if self.cfg.apply_bounds_shim:
                _, _, _, h, w = batch["context"]["image"].shape
                batch = apply_bounds_shim(batch, self.cfg.near_disparity * min(h, w), 0.5)
            return batch
        return data_shim

    @property
    def sampler(self):
        return self.epipolar_transformer.epipolar_sampler

    @property
    def last_layer_weights(self) -> Tensor:
        return self.to_gaussians[-1].weight
