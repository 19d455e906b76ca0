"""KL-f8 VAE network with the parameter tree of `diffusers.AutoencoderKL` (diffusers==0.25.1, the reference's
pinned dependency: requirements.txt:12, autoencoder_kl.py:5-12, 48-57).

diffusers is not installed in this image, so the public architecture is restated from its definition
(`AutoencoderKL`, `Encoder`/`Decoder`, `UNetMidBlock2D`, `Down/UpDecoderBlock2D`, `ResnetBlock2D`, deprecated
single-head `Attention` block) with IDENTICAL module/parameter names, so that `pretrained/autoencoder/kl_f8.pt`
and latentSplat checkpoints (`autoencoder.model.decoder.up_blocks...`) load with strict=True.
PARITY UNPINNED against diffusers itself; tests check the parameter inventory (names + shapes of kl-f8:
83 653 863 parameters) and module semantics against independent torch re-computation.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F
from torch import Tensor, nn
from latentsplat_b200 import fmha  # attention cores on our kernels (CUDA)
from latentsplat_b200.gemm import Linear  # nn.Linear with tcgen05 TF32 GEMMs on CUDA
from latentsplat_b200.norm import GroupNorm, group_norm_tokens  # GroupNorm with the following SiLU fused in (sm_100a kernels on CUDA)
from latentsplat_b200.conv import Conv2d  # nn.Conv2d on the tcgen05 implicit-GEMM kernels (NHWC / channels_last on CUDA)


class ResnetBlock2D(nn.Module):
    def __init__(self, in_channels: int, out_channels: int, groups: int = 32, eps: float = 1e-6):
        super().__init__()
        self.norm1 = GroupNorm(groups, in_channels, eps=eps, affine=True, act="silu")      # norm + nonlinearity
        self.conv1 = Conv2d(in_channels, out_channels, 3, 1, 1)
        self.norm2 = GroupNorm(groups, out_channels, eps=eps, affine=True, act="silu")
        self.dropout = nn.Dropout(0.0)
        self.conv2 = Conv2d(out_channels, out_channels, 3, 1, 1)
        self.nonlinearity = nn.SiLU()
        self.conv_shortcut = Conv2d(in_channels, out_channels, 1, 1, 0) if in_channels != out_channels else None

    def forward(self, x: Tensor) -> Tensor:
        h = self.conv1(self.norm1(x))                       # diffusers: conv1(nonlinearity(norm1(x))), SiLU fused into the norm
        h = self.conv2(self.dropout(self.norm2(h)))
        if self.conv_shortcut is not None:
            x = self.conv_shortcut(x)
        return x + h


class Attention(nn.Module):
    """Single-head spatial self-attention with residual (the VAE mid-block attention of diffusers)."""

    def __init__(self, channels: int, groups: int = 32, eps: float = 1e-6):
        super().__init__()
        self.group_norm = GroupNorm(groups, channels, eps=eps, affine=True)
        self.to_q = Linear(channels, channels)
        self.to_k = Linear(channels, channels)
        self.to_v = Linear(channels, channels)
        self.to_out = nn.ModuleList([Linear(channels, channels), nn.Dropout(0.0)])

    def forward(self, x: Tensor) -> Tensor:
        b, c, h, w = x.shape
        # token-major (b, hw, c) view: free for channels_last activations (the convolutions' layout), one copy for NCHW
        t = x.permute(0, 2, 3, 1).reshape(b, h * w, c)
        gn = self.group_norm
        t = group_norm_tokens(t, gn.num_groups, gn.weight, gn.bias, gn.eps)
        q, k, v = self.to_q(t), self.to_k(t), self.to_v(t)
        if fmha.wide_supported(q, k, v):
            t = fmha.attention_wide(q, k, v, c ** -0.5)        # scores through our GEMM + in-place row softmax
        else:
            t = F.scaled_dot_product_attention(q[:, None], k[:, None], v[:, None])[:, 0]
        t = self.to_out[1](self.to_out[0](t))
        # x first: the sum inherits x's memory format, so the residual stream keeps one layout up the decoder
        return x + t.reshape(b, h, w, c).permute(0, 3, 1, 2)


class UNetMidBlock2D(nn.Module):
    def __init__(self, channels: int, groups: int = 32, eps: float = 1e-6):
        super().__init__()
        self.attentions = nn.ModuleList([Attention(channels, groups, eps)])
        self.resnets = nn.ModuleList([ResnetBlock2D(channels, channels, groups, eps),
                                      ResnetBlock2D(channels, channels, groups, eps)])

    def forward(self, x: Tensor) -> Tensor:
        return self.resnets[1](self.attentions[0](self.resnets[0](x)))


class Downsample2D(nn.Module):
    def __init__(self, channels: int):
        super().__init__()
        self.conv = Conv2d(channels, channels, 3, stride=2, padding=0)

    def forward(self, x: Tensor) -> Tensor:
        return self.conv(F.pad(x, (0, 1, 0, 1), mode="constant", value=0))


class Upsample2D(nn.Module):
    def __init__(self, channels: int):
        super().__init__()
        self.conv = Conv2d(channels, channels, 3, padding=1)

    def forward(self, x: Tensor) -> Tensor:
        # conv(F.interpolate(x, scale_factor=2.0, mode="nearest")), with the up-sampling folded into the filter on CUDA
        return self.conv.forward_upsampled2x(x)


class DownEncoderBlock2D(nn.Module):
    def __init__(self, in_channels: int, out_channels: int, num_layers: int, add_downsample: bool, groups: int = 32):
        super().__init__()
        self.resnets = nn.ModuleList([ResnetBlock2D(in_channels if i == 0 else out_channels, out_channels, groups)
                                      for i in range(num_layers)])
        self.downsamplers = nn.ModuleList([Downsample2D(out_channels)]) if add_downsample else None

    def forward(self, x: Tensor) -> Tensor:
        for r in self.resnets:
            x = r(x)
        if self.downsamplers is not None:
            x = self.downsamplers[0](x)
        return x


class UpDecoderBlock2D(nn.Module):
    def __init__(self, in_channels: int, out_channels: int, num_layers: int, add_upsample: bool, groups: int = 32):
        super().__init__()
        self.resnets = nn.ModuleList([ResnetBlock2D(in_channels if i == 0 else out_channels, out_channels, groups)
                                      for i in range(num_layers)])
        self.upsamplers = nn.ModuleList([Upsample2D(out_channels)]) if add_upsample else None

    def forward(self, x: Tensor) -> Tensor:
        for r in self.resnets:
            x = r(x)
        if self.upsamplers is not None:
            x = self.upsamplers[0](x)
        return x


class Encoder(nn.Module):
    def __init__(self, in_channels: int, latent_channels: int, block_out_channels, layers_per_block: int, groups: int = 32):
        super().__init__()
        self.conv_in = Conv2d(in_channels, block_out_channels[0], 3, 1, 1)
        self.down_blocks = nn.ModuleList()
        out = block_out_channels[0]
        for i, ch in enumerate(block_out_channels):
            inp, out = out, ch
            self.down_blocks.append(DownEncoderBlock2D(inp, out, layers_per_block, i != len(block_out_channels) - 1, groups))
        self.mid_block = UNetMidBlock2D(block_out_channels[-1], groups)
        self.conv_norm_out = GroupNorm(groups, block_out_channels[-1], eps=1e-6, act="silu")
        self.conv_act = nn.SiLU()
        self.conv_out = Conv2d(block_out_channels[-1], 2 * latent_channels, 3, padding=1)

    def forward(self, x: Tensor) -> Tensor:
        x = self.conv_in(x)
        for blk in self.down_blocks:
            x = blk(x)
        x = self.mid_block(x)
        return self.conv_out(self.conv_norm_out(x))        # conv_act (SiLU) is fused into conv_norm_out


class Decoder(nn.Module):
    def __init__(self, latent_channels: int, out_channels: int, block_out_channels, layers_per_block: int, groups: int = 32):
        super().__init__()
        self.conv_in = Conv2d(latent_channels, block_out_channels[-1], 3, 1, 1)
        self.mid_block = UNetMidBlock2D(block_out_channels[-1], groups)
        self.up_blocks = nn.ModuleList()
        rev = list(reversed(block_out_channels))
        out = rev[0]
        for i, ch in enumerate(rev):
            inp, out = out, ch
            self.up_blocks.append(UpDecoderBlock2D(inp, out, layers_per_block + 1, i != len(rev) - 1, groups))
        self.conv_norm_out = GroupNorm(groups, block_out_channels[0], eps=1e-6, act="silu")
        self.conv_act = nn.SiLU()
        self.conv_out = Conv2d(block_out_channels[0], out_channels, 3, padding=1)

    def forward(self, z: Tensor) -> Tensor:
        z = self.mid_block(self.conv_in(z))
        for blk in self.up_blocks:
            z = blk(z)
        return self.conv_out(self.conv_norm_out(z))        # conv_act (SiLU) is fused into conv_norm_out


class AutoencoderKLModel(nn.Module):
    """`diffusers.AutoencoderKL(in, out, down_block_types, up_block_types, block_out_channels, layers_per_block,
    latent_channels, sample_size)` restricted to DownEncoderBlock2D / UpDecoderBlock2D."""

    def __init__(self, in_channels: int = 3, out_channels: int = 3, block_out_channels=(128, 256, 512, 512),
                 layers_per_block: int = 2, latent_channels: int = 4, norm_num_groups: int = 32):
        super().__init__()
        self.encoder = Encoder(in_channels, latent_channels, block_out_channels, layers_per_block, norm_num_groups)
        self.decoder = Decoder(latent_channels, out_channels, block_out_channels, layers_per_block, norm_num_groups)
        self.quant_conv = Conv2d(2 * latent_channels, 2 * latent_channels, 1)
        self.post_quant_conv = Conv2d(latent_channels, latent_channels, 1)

    def encode_moments(self, x: Tensor) -> tuple[Tensor, Tensor]:
        mean, logvar = self.quant_conv(self.encoder(x)).chunk(2, dim=1)
        return mean, torch.clamp(logvar, -30.0, 20.0)

    def decode(self, z: Tensor) -> Tensor:
        return self.decoder(self.post_quant_conv(z))
