"""DINO ViT backbone + token MLPs (/root/reference/src/model/encoder/backbone/backbone_dino.py:14-88).

Outputs are identical to the reference's; the data movement is not: the reference materialises the 8x
nearest-neighbour `repeat` of the local tokens at full resolution ((bv, 512, 256, 256) fp32 = 1.07 GB at the
training shape) and the encoder then applies ReLU + Linear(512 -> 128) per *pixel*
(encoder_epipolar.py:143-145).  `forward_tokens` exposes the un-repeated (bv, d_out, h/8, w/8) map so that the
encoder can apply that pointwise projection before the repeat (64x fewer FLOPs and bytes, algebraically
identical because ReLU and Linear are pointwise and `repeat` only replicates).
"""
from __future__ import annotations

import os
import warnings
from dataclasses import dataclass
from fractions import Fraction
from pathlib import Path
from typing import Literal, Optional

import torch

import torch.nn.functional as F
from torch import Tensor, nn

from .backbone import Backbone
from .dino_vit import CONFIGS, build_dino
from latentsplat_b200.gemm import Linear  # nn.Linear with tcgen05 TF32 GEMMs on CUDA


@dataclass
class BackboneDinoCfg:
    name: Literal["dino"]
    model: Literal["dino_vits16", "dino_vits8", "dino_vitb16", "dino_vitb8"]
    upscale_mode: Literal["interpolate", "repeat"] = "repeat"
    # The reference obtains PRETRAINED weights: torch.hub.load("facebookresearch/dino:main", cfg.model) (backbone_dino.py:33).
    # There is no network here, so the official state_dict (same parameter names as our restatement, loaded strict=True) comes
    # from a file: this field, else $LS_DINO_WEIGHTS, else pretrained/backbone/<model>.pth.  "random" = deliberately untrained
    # (benchmarks, tests); None with no file present warns loudly once per construction.
    pretrained: Optional[str] = None


class BackboneDino(Backbone[BackboneDinoCfg]):
    def __init__(self, cfg: BackboneDinoCfg, d_in: int, d_out: int, scale_factor: Fraction) -> None:
        super().__init__(cfg, d_in, d_out, scale_factor)
        assert d_in == 3
        if self.cfg.upscale_mode == "repeat":
            n = self.patch_size * Fraction(self.scale_factor)
            assert n.denominator == 1
            self.n_repeats = int(n)
        self.dino = build_dino(cfg.model)
        self._load_pretrained()
        d = CONFIGS[cfg.model][1]
        # NB: the reference hard-codes 768 (ViT-B); ViT-S checkpoints would not fit it either.
        # Linear, ReLU, Linear (:34-43); the ReLU runs in the first GEMM's epilogue, an Identity keeps the Sequential indices
        self.global_token_mlp = nn.Sequential(Linear(d, d, act="relu"), nn.Identity(), Linear(d, self.d_out))
        self.local_token_mlp = nn.Sequential(Linear(d, d, act="relu"), nn.Identity(), Linear(d, self.d_out))

    def _load_pretrained(self) -> None:
        choice = self.cfg.pretrained or os.environ.get("LS_DINO_WEIGHTS")
        if choice == "random":
            return
        path = Path(choice) if choice else Path("pretrained") / "backbone" / f"{self.cfg.model}.pth"
        if not path.exists():
            if choice:
                raise FileNotFoundError(f"DINO weights {path} not found (BackboneDinoCfg.pretrained / LS_DINO_WEIGHTS)")
            warnings.warn(f"BackboneDino: no pretrained weights at {path}; the reference starts from torch.hub's pretrained "
                          f"{self.cfg.model} -- this backbone is RANDOMLY initialised (set pretrained='random' to silence)",
                          stacklevel=3)
            return
        state = torch.load(path, map_location="cpu", weights_only=True)
        self.dino.load_state_dict(state.get("state_dict", state), strict=True)

    @property
    def patch_size(self) -> int:
        return int("".join(filter(str.isdigit, self.cfg.model)))

    def forward_tokens(self, x: Tensor) -> tuple[Tensor, Tensor]:
        """-> local (b, d_out, h/ps, w/ps) and global (b, d_out, 1, 1) token features, before any upscaling."""
        b, _, h, w = x.shape
        assert h % self.patch_size == 0 and w % self.patch_size == 0
        tokens = self.dino.get_intermediate_layers(x)[0]
        global_token = self.global_token_mlp(tokens[:, 0]).view(b, -1, 1, 1)
        local_tokens = self.local_token_mlp(tokens[:, 1:])
        local_tokens = local_tokens.transpose(1, 2).unflatten(2, (h // self.patch_size, w // self.patch_size))
        return local_tokens, global_token

    def forward(self, x: Tensor) -> Tensor:
        b, _, h, w = x.shape
        local_tokens, global_token = self.forward_tokens(x)
        if self.cfg.upscale_mode == "interpolate":
            size = tuple(int(Fraction(self.scale_factor) * s) for s in (h, w))
            local_tokens = F.interpolate(local_tokens, size, mode="bilinear", align_corners=True)
        elif self.cfg.upscale_mode == "repeat":
            local_tokens = local_tokens.repeat_interleave(self.n_repeats, dim=2).repeat_interleave(self.n_repeats, dim=3)
        else:
            raise ValueError(f"Unknown upscale_mode {self.cfg.upscale_mode}")
        return local_tokens + global_token
