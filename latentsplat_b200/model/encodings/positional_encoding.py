"""Sinusoidal encoding of values in [0,1] over octaves (/root/reference/src/model/encodings/positional_encoding.py:8-36)."""
from __future__ import annotations

import torch
import torch.nn as nn
from torch import Tensor


class PositionalEncoding(nn.Module):
    frequencies: Tensor  # (octave, 2)
    phases: Tensor       # (octave, 2)

    def __init__(self, num_octaves: int):
        super().__init__()
        octaves = torch.arange(num_octaves).float()
        frequencies = (2 * torch.pi * 2 ** octaves)[:, None].repeat(1, 2)   # lowest frequency: period 1
        self.register_buffer("frequencies", frequencies, persistent=False)
        phases = torch.tensor([0, 0.5 * torch.pi], dtype=torch.float32)[None].repeat(num_octaves, 1)  # sin, cos
        self.register_buffer("phases", phases, persistent=False)

    def forward(self, samples: Tensor) -> Tensor:
        """(…, dim) -> (…, dim * octaves * 2), ordered (dim, octave, phase)."""
        x = samples[..., None, None] * self.frequencies
        return torch.sin(x + self.phases).flatten(-3)

    def d_out(self, dimensionality: int) -> int:
        return self.frequencies.numel() * dimensionality
