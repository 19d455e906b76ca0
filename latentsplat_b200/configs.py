"""Typed configurations of the shipped experiments, restated from the reference's YAML
(config/model/encoder/epipolar.yaml, config/model/encoder/backbone/dino.yaml,
config/model/autoencoder/kl_f8_skip.yaml, config/experiment/re10k.yaml) -- no hydra needed."""
from __future__ import annotations

from fractions import Fraction

from .model.autoencoder import AutoencoderKLCfg, get_autoencoder
from .model.decoder import DecoderSplattingCUDACfg, get_decoder
from .model.discriminator import DiscriminatorPatchGanCfg, get_discriminator
from .model.encoder import get_encoder
from .model.encoder.backbone.backbone_dino import BackboneDinoCfg
from .model.encoder.common.gaussian_adapter import GaussianAdapterCfg
from .model.encoder.encoder_epipolar import EncoderEpipolarCfg, OpacityMappingCfg
from .model.encoder.epipolar.epipolar_transformer import EpipolarTransformerCfg
from .model.encoder.epipolar.image_self_attention import ImageSelfAttentionCfg


def encoder_cfg() -> EncoderEpipolarCfg:
    """config/model/encoder/epipolar.yaml with the re10k / co3d experiment overrides."""
    return EncoderEpipolarCfg(
        name="epipolar", d_backbone=512, d_feature=128, num_monocular_samples=32, num_surfaces=1, predict_opacity=False,
        backbone=BackboneDinoCfg("dino", "dino_vitb8", pretrained="random"), near_disparity=3.0,
        gaussian_adapter=GaussianAdapterCfg(gaussian_scale_min=0.5, gaussian_scale_max=15.0, color_sh_degree=4,
                                            feature_sh_degree=2),
        apply_bounds_shim=True,
        epipolar_transformer=EpipolarTransformerCfg(
            self_attention=ImageSelfAttentionCfg(patch_size=4, num_octaves=10, num_layers=2, num_heads=4, d_token=128,
                                                 d_dot=128, d_mlp=256),
            num_octaves=10, num_layers=2, num_heads=4, num_samples=32, d_dot=128, d_mlp=256, downscale=4),
        opacity_mapping=OpacityMappingCfg(initial=0.0, final=0.0, warm_up=1),
        gaussians_per_pixel=3, use_epipolar_transformer=True, use_transmittance=False)


def autoencoder_cfg(pretrained: bool = False) -> AutoencoderKLCfg:
    """config/model/autoencoder/kl_f8_skip.yaml."""
    return AutoencoderKLCfg(name="kl", model="kl_f8", down_block_types=["DownEncoderBlock2D"] * 4,
                            up_block_types=["UpDecoderBlock2D"] * 4, block_out_channels=[128, 256, 512, 512],
                            layers_per_block=2, latent_channels=4, skip_connections=True, skip_extra=True,
                            skip_zero=True, pretrained=pretrained)


def build_modules(variational: str = "gaussians", supersampling_factor: int = 8, background_color=(0.0, 0.0, 0.0),
                  pretrained: bool = False, with_discriminator: bool = True):
    """The module construction of /root/reference/src/main.py:107-129 (encode_latents=false)."""
    autoencoder = get_autoencoder(autoencoder_cfg(pretrained))
    encoder, _ = get_encoder(encoder_cfg(), d_in=3, n_feature_channels=autoencoder.d_latent,
                             scale_factor=Fraction(supersampling_factor, autoencoder.downscale_factor),
                             variational=variational != "none")
    decoder = get_decoder(DecoderSplattingCUDACfg("splatting_cuda"), list(background_color), variational == "latents")
    disc = get_discriminator(DiscriminatorPatchGanCfg("patch_gan", "kl_f8", pretrained=pretrained)) \
        if with_discriminator else None
    return autoencoder, encoder, decoder, disc
