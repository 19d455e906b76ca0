"""Index tables for "every view against every OTHER view" (/root/reference/src/misc/heterogeneous_pairings.py)."""
from __future__ import annotations

import torch
from torch import Tensor


def generate_heterogeneous_index(n: int, device=torch.device("cpu")) -> tuple[Tensor, Tensor]:
    """(self, other) indices of shape (n, n-1): row i lists every j != i in ascending order (:9-24)."""
    arange = torch.arange(n, device=device)
    index_self = arange[:, None].expand(n, n - 1)
    index_other = arange[None, :].repeat(n, 1) + torch.ones((n, n), device=device, dtype=torch.int64).triu()
    return index_self, index_other[:, :-1]


def generate_heterogeneous_index_transpose(n: int, device=torch.device("cpu")) -> tuple[Tensor, Tensor]:
    """Index pair that swaps the roles of view and other-view; an involution (:27-43)."""
    arange = torch.arange(n, device=device)
    upper = torch.ones((n, n), device=device, dtype=torch.int64).triu()
    index_self = arange[None, :].repeat(n, 1) + upper
    index_other = arange[:, None].expand(n, n) - (1 - upper)
    return index_self[:, :-1], index_other[:, :-1]
