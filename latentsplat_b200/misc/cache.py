"""Tiny per-device cache of constant tensors.  `torch.tensor(python_list, device=cuda)` is a pageable H2D copy:
a stream synchronisation on every call and illegal during CUDA-graph capture."""
from __future__ import annotations

import torch

_CACHE: dict = {}


def device_constant(values, device, dtype=torch.float32) -> torch.Tensor:
    key = (tuple(values), str(device), dtype)
    if key not in _CACHE:
        _CACHE[key] = torch.tensor(tuple(values), dtype=dtype, device=device)
    return _CACHE[key]
