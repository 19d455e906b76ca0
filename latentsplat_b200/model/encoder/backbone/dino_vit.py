"""DINO ViT (facebookresearch/dino `vision_transformer.py`, `dino_vitb8` = patch 8, dim 768, depth 12, heads 12).

The reference obtains this network with `torch.hub.load("facebookresearch/dino:main", cfg.model)`
(backbone_dino.py:33: network access, un-pinned branch) -- not available here, so the public architecture is
restated with the SAME parameter names (`cls_token`, `pos_embed`, `patch_embed.proj`, `blocks.N.norm1`,
`blocks.N.attn.qkv`, `blocks.N.attn.proj`, `blocks.N.norm2`, `blocks.N.mlp.fc1/fc2`, `norm`) so that DINO / latentSplat
checkpoints load.  PARITY UNPINNED against the hub model (cannot be downloaded here).
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F
from torch import Tensor, nn
from latentsplat_b200 import fmha  # tcgen05 flash-attention core (CUDA)
from latentsplat_b200.gemm import Linear  # nn.Linear with tcgen05 TF32 GEMMs on CUDA
from latentsplat_b200.norm import LayerNorm  # nn.LayerNorm on our warp-per-row kernel on CUDA
from latentsplat_b200.conv import Conv2d  # nn.Conv2d with the bias add / bias gradient on our kernels (CUDA)


class Mlp(nn.Module):
    def __init__(self, dim: int, hidden: int):
        super().__init__()
        self.fc1 = Linear(dim, hidden, act="gelu")          # GELU in the GEMM epilogue (pre-activation kept for backward)
        self.act = nn.Identity()                             # upstream: nn.GELU() here (no parameters: checkpoint-neutral)
        self.fc2 = Linear(hidden, dim)

    def forward(self, x, residual=None):
        """fc2(gelu(fc1(x))) (+ residual): the block's skip connection rides in fc2's epilogue."""
        return self.fc2(self.act(self.fc1(x)), residual=residual)


class _Bf16AttentionCore(torch.autograd.Function):
    """softmax(q k^T scale) v on ONE bf16 copy of the packed fp32 qkv projection, through the library flash kernel.

    Keeps the casts / layout shuffles around the library call to the minimum: forward = one fp32->bf16 cast of
    (B, N, 3C) + one bf16->fp32 cast of the output; backward = one cast of the incoming gradient + three strided
    cast-copies of dq, dk, dv straight into the packed (B, N, 3, H, D) gradient (autograd's own route through
    permute / unbind / .to() costs a zero-fill and an add per tensor on top of that)."""

    @staticmethod
    def forward(ctx, qkv: Tensor, heads: int, scale: float) -> Tensor:
        B, N, C3 = qkv.shape
        D = C3 // 3 // heads
        packed = qkv.to(torch.bfloat16).view(B, N, 3, heads, D)
        q, k, v = (packed[:, :, i].transpose(1, 2) for i in range(3))                 # (B, H, N, D) strided views
        if ctx.needs_input_grad[0]:
            q, k, v = (t.detach().requires_grad_(True) for t in (q, k, v))
            with torch.enable_grad():
                out = F.scaled_dot_product_attention(q, k, v, scale=scale)
            ctx.inner = (q, k, v, out)
        else:
            out = F.scaled_dot_product_attention(q, k, v, scale=scale)
        return out.detach().transpose(1, 2).reshape(B, N, heads * D).to(qkv.dtype)

    @staticmethod
    def backward(ctx, grad: Tensor):
        q, k, v, out = ctx.inner              # kept (not cleared): a second backward through a retained graph must still work
        B, H, N, D = q.shape
        g = grad.to(torch.bfloat16).view(B, N, H, D).transpose(1, 2)
        dq, dk, dv = torch.autograd.grad(out, (q, k, v), g, retain_graph=True)
        packed = torch.empty((B, N, 3, H, D), dtype=grad.dtype, device=grad.device)
        for i, d in enumerate((dq, dk, dv)):
            packed[:, :, i].copy_(d.transpose(1, 2))
        return packed.view(B, N, 3 * H * D), None, None


class Attention(nn.Module):
    def __init__(self, dim: int, num_heads: int):
        super().__init__()
        self.num_heads = num_heads
        self.scale = (dim // num_heads) ** -0.5
        self.qkv = Linear(dim, dim * 3, bias=True)
        self.proj = Linear(dim, dim)

    def forward(self, x, residual=None):
        """proj(attention(qkv(x))) (+ residual): the block's skip connection rides in proj's epilogue."""
        B, N, C = x.shape
        qkv = self.qkv(x)
        if fmha.supported(qkv, self.num_heads):
            # our tcgen05 flash-attention core, reading the packed fp32 projection in place (TF32 operands)
            return self.proj(fmha.attention_packed(qkv, self.num_heads, self.scale), residual=residual)
        if x.is_cuda and ATTENTION_BF16:
            # A/B arm (fmha.ENABLED = False): library flash-attention on a bf16 copy of q/k/v
            return self.proj(_Bf16AttentionCore.apply(qkv, self.num_heads, self.scale), residual=residual)
        qkv = qkv.reshape(B, N, 3, self.num_heads, C // self.num_heads).permute(2, 0, 3, 1, 4)
        x = F.scaled_dot_product_attention(qkv[0], qkv[1], qkv[2], scale=self.scale)
        return self.proj(x.transpose(1, 2).reshape(B, N, C), residual=residual)


class Block(nn.Module):
    def __init__(self, dim: int, num_heads: int, mlp_ratio: float = 4.0):
        super().__init__()
        self.norm1 = LayerNorm(dim, eps=1e-6)
        self.attn = Attention(dim, num_heads)
        self.norm2 = LayerNorm(dim, eps=1e-6)
        self.mlp = Mlp(dim, int(dim * mlp_ratio))

    def forward(self, x):
        x = self.attn(self.norm1(x), residual=x)             # x + attn(norm1(x))
        return self.mlp(self.norm2(x), residual=x)           # x + mlp(norm2(x))


class PatchEmbed(nn.Module):
    def __init__(self, img_size: int, patch_size: int, in_chans: int, embed_dim: int):
        super().__init__()
        self.num_patches = (img_size // patch_size) ** 2
        self.proj = Conv2d(in_chans, embed_dim, kernel_size=patch_size, stride=patch_size)

    def forward(self, x):
        return self.proj(x).flatten(2).transpose(1, 2)


ATTENTION_BF16 = True     # comparator only: library bf16 flash-attention, used when latentsplat_b200.fmha.ENABLED is False

CONFIGS = {"dino_vits16": (16, 384, 12, 6), "dino_vits8": (8, 384, 12, 6),
           "dino_vitb16": (16, 768, 12, 12), "dino_vitb8": (8, 768, 12, 12)}


class VisionTransformer(nn.Module):
    def __init__(self, patch_size: int = 8, embed_dim: int = 768, depth: int = 12, num_heads: int = 12,
                 img_size: int = 224, in_chans: int = 3):
        super().__init__()
        self.patch_size = patch_size
        self.embed_dim = embed_dim
        self.patch_embed = PatchEmbed(img_size, patch_size, in_chans, embed_dim)
        self.cls_token = nn.Parameter(torch.zeros(1, 1, embed_dim))
        self.pos_embed = nn.Parameter(torch.zeros(1, self.patch_embed.num_patches + 1, embed_dim))
        self.blocks = nn.ModuleList([Block(embed_dim, num_heads) for _ in range(depth)])
        self.norm = LayerNorm(embed_dim, eps=1e-6)
        self._pos_cache: dict = {}
        nn.init.trunc_normal_(self.pos_embed, std=0.02)
        nn.init.trunc_normal_(self.cls_token, std=0.02)

    def _bicubic_matrix(self, side: int, w0: float, h0: float, device) -> Tensor:
        """Upstream resamples the patch position embeddings with `F.interpolate(..., scale_factor=(w0/side, h0/side),
        mode="bicubic")`.  That map is linear and depends only on the sizes, so it is captured once as a
        (side^2, out^2) matrix (by pushing an identity basis through the very same call) and applied as a GEMM:
        the CUDA bicubic kernel costs 6.3 ms forward + 1.8 ms backward per step on a (1,768,28,28) input."""
        key = (side, round(w0, 3), round(h0, 3), str(device))
        if key not in self._pos_cache:
            basis = torch.eye(side * side, dtype=torch.float32).reshape(1, side * side, side, side)
            out = F.interpolate(basis, scale_factor=(w0 / side, h0 / side), mode="bicubic")
            assert int(w0) == out.shape[-2] and int(h0) == out.shape[-1]
            self._pos_cache[key] = out.flatten(2)[0].to(device)              # (side^2, out_h * out_w)
        return self._pos_cache[key]

    def interpolate_pos_encoding(self, x: Tensor, w: int, h: int) -> Tensor:
        npatch = x.shape[1] - 1
        N = self.pos_embed.shape[1] - 1
        if npatch == N and w == h:
            return self.pos_embed
        class_pos, patch_pos = self.pos_embed[:, 0], self.pos_embed[:, 1:]
        w0, h0 = w // self.patch_size + 0.1, h // self.patch_size + 0.1   # +0.1: upstream's guard against rounding
        side = int(math.sqrt(N))
        m = self._bicubic_matrix(side, w0, h0, x.device)
        patch_pos = (m.t() @ patch_pos[0])[None]                            # (1, out^2, dim)
        return torch.cat((class_pos.unsqueeze(0), patch_pos), dim=1)

    def prepare_tokens(self, x: Tensor) -> Tensor:
        B, _, w, h = x.shape
        x = self.patch_embed(x)
        x = torch.cat((self.cls_token.expand(B, -1, -1), x), dim=1)
        return x + self.interpolate_pos_encoding(x, w, h)

    def get_intermediate_layers(self, x: Tensor, n: int = 1) -> list[Tensor]:
        """Normalised outputs of the last n blocks."""
        x = self.prepare_tokens(x)
        output = []
        for i, blk in enumerate(self.blocks):
            x = blk(x)
            if len(self.blocks) - i <= n:
                output.append(self.norm(x))
        return output

    def forward(self, x: Tensor) -> Tensor:
        return self.get_intermediate_layers(x)[0][:, 0]


def build_dino(model: str) -> VisionTransformer:
    patch, dim, depth, heads = CONFIGS[model]
    return VisionTransformer(patch, dim, depth, heads)
