"""One switch for the arithmetic of the contractions on the hot path.

The reference runs its Linear layers and attention as fp32 SIMT GEMMs (torch default matmul precision "highest") and its
convolutions as cuDNN TF32 (torch default `cudnn.allow_tf32 = True`).  This package's default routes every Linear, every
convolution and the attention cores through tcgen05 `kind::tf32` kernels (fp32 in HBM, 10-bit-mantissa operands in the tensor
core, fp32 accumulation in TMEM): the tensor core has no fp32 MMA, so an sm_100a-native path is TF32 by construction.

    set_math_mode("tf32")   default: our tcgen05 kernels everywhere
    set_math_mode("fp32")   parity debugging: the contraction kernels of ours are switched off and torch's library ops run them in
                            full fp32 (cuBLAS SIMT GEMMs, cuDNN with TF32 disabled, fp32 SDPA); the norm / sampler / rasterizer
                            kernels are fp32 either way and stay on

`LS_MATH_MODE=tf32|fp32` in the environment selects the mode at import; bench.py records it in `config.precision`.
"""
from __future__ import annotations

import os

import torch

_MODE = "tf32"


def set_math_mode(mode: str) -> None:
    global _MODE
    if mode not in ("tf32", "fp32"):
        raise ValueError(f"math mode must be 'tf32' or 'fp32', got {mode!r}")
    from . import conv, fmha, gemm
    on = mode == "tf32"
    gemm.enabled = conv.ENABLED = fmha.ENABLED = on
    from .model.encoder.backbone import dino_vit
    dino_vit.ATTENTION_BF16 = False if not on else dino_vit.ATTENTION_BF16       # never bf16 in the fp32 mode
    torch.backends.cuda.matmul.allow_tf32 = False                               # library GEMMs stay fp32 in both modes
    torch.backends.cudnn.allow_tf32 = on
    _MODE = mode


def math_mode() -> str:
    return _MODE


if os.environ.get("LS_MATH_MODE"):
    set_math_mode(os.environ["LS_MATH_MODE"])
