"""LPIPS (VGG-16) perceptual loss (/root/reference/src/loss/loss_lpips.py:19-60: `lpips.LPIPS(net="vgg")` with
`normalize=True`, mean over images; parameters converted to non-persistent buffers, i.e. frozen and absent from checkpoints).

The `lpips` package is not installed in this image, so the public definition (richzhang/PerceptualSimilarity v0.1) is restated:
inputs in [0, 1] -> 2x - 1 -> channel-wise (x - shift) / scale -> VGG-16 features after relu1_2, relu2_2, relu3_3, relu4_3, relu5_3
-> unit-normalise every pixel's feature vector -> squared difference -> non-negative 1x1 `lin` weights -> spatial mean -> sum of the
five layers.  The 13 convolutions run on our tcgen05 implicit-GEMM kernels on CUDA (bias + ReLU in the epilogue; ~15.4 GMAC per
256x256 image and side, forward + input gradient).  PARITY UNPINNED against the package and its weights: without a weight file the
network is randomly initialised and a warning says so (`LS_LPIPS_WEIGHTS` / `pretrained/lpips_vgg.pth`: a state dict holding
torchvision's `features.N.{weight,bias}` and lpips' `lin{k}.model.1.weight`)."""
from __future__ import annotations

import os
import warnings
from dataclasses import dataclass
from pathlib import Path
from typing import Literal, Optional

import torch
import torch.nn.functional as F
from torch import Tensor, nn

from ..conv import Conv2d
from .loss import Loss, LossCfg

# torchvision.models.vgg16().features indices of the convolutions, grouped by the five LPIPS slices (a max-pool opens slices 2..5)
_SLICES = (((0, 3, 64), (2, 64, 64)),
           ((5, 64, 128), (7, 128, 128)),
           ((10, 128, 256), (12, 256, 256), (14, 256, 256)),
           ((17, 256, 512), (19, 512, 512), (21, 512, 512)),
           ((24, 512, 512), (26, 512, 512), (28, 512, 512)))
_CHANNELS = (64, 128, 256, 512, 512)


@dataclass
class LossLpipsCfg(LossCfg):
    name: Literal["lpips"] = "lpips"


class LpipsVgg(nn.Module):
    def __init__(self, weights: Optional[str] = None) -> None:
        super().__init__()
        self.register_buffer("shift", torch.tensor([-0.030, -0.088, -0.188]).view(1, 3, 1, 1), persistent=False)
        self.register_buffer("scale", torch.tensor([0.458, 0.448, 0.450]).view(1, 3, 1, 1), persistent=False)
        self.slices = nn.ModuleList(nn.ModuleList(Conv2d(cin, cout, 3, padding=1, act="relu") for _, cin, cout in group)
                                    for group in _SLICES)
        self.lins = nn.ParameterList(nn.Parameter(torch.rand(1, c, 1, 1) / c) for c in _CHANNELS)     # non-negative
        self._load(weights)
        for p in self.parameters():            # frozen, as the reference's convert_to_buffer (loss_lpips.py:27)
            p.requires_grad_(False)

    def _load(self, weights: Optional[str]) -> None:
        choice = weights or os.environ.get("LS_LPIPS_WEIGHTS")
        if choice == "random":
            return
        path = Path(choice) if choice else Path("pretrained") / "lpips_vgg.pth"
        if not path.exists():
            if choice:
                raise FileNotFoundError(f"LPIPS weights {path} not found (LS_LPIPS_WEIGHTS)")
            warnings.warn(f"LpipsVgg: no weights at {path}; the reference uses the pretrained lpips / VGG-16 weights -- this network is "
                          "RANDOMLY initialised (pass weights='random' to silence)", stacklevel=3)
            return
        state = torch.load(path, map_location="cpu", weights_only=True)
        with torch.no_grad():
            for group, convs in zip(_SLICES, self.slices):
                for (idx, _, _), conv in zip(group, convs):
                    conv.weight.copy_(state[f"features.{idx}.weight"])
                    conv.bias.copy_(state[f"features.{idx}.bias"])
            for k, lin in enumerate(self.lins):
                lin.copy_(state[f"lin{k}.model.1.weight"])

    def features(self, x: Tensor):
        out = []
        for k, convs in enumerate(self.slices):
            if k:
                x = F.max_pool2d(x, 2, 2)
            for conv in convs:
                x = conv(x)                    # bias + ReLU in the convolution's epilogue on CUDA
            out.append(x)
        return out

    def forward(self, pred: Tensor, target: Tensor) -> Tensor:
        """pred, target (n, 3, h, w) in [0, 1] -> mean LPIPS distance."""
        n = pred.shape[0]
        both = (2.0 * torch.cat((pred, target)) - 1.0 - self.shift) / self.scale           # one pass through VGG for both sides
        total = 0.0
        for feat, lin in zip(self.features(both), self.lins):
            feat = feat / (feat.square().sum(dim=1, keepdim=True).sqrt() + 1e-10)
            diff = (feat[:n] - feat[n:]).square()
            total = total + (diff * lin).sum(dim=1).mean(dim=(1, 2))
        return total.mean()


class LossLpips(Loss):
    def __init__(self, cfg: LossLpipsCfg, lpips: Optional[LpipsVgg] = None) -> None:
        super().__init__(cfg)
        self.lpips = LpipsVgg() if lpips is None else lpips

    def unweighted_loss(self, prediction, gt) -> Tensor:
        return self.lpips(prediction.image.flatten(0, 1), gt.image.flatten(0, 1))
