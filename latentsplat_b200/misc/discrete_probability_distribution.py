"""Sampling of depth buckets (/root/reference/src/misc/discrete_probability_distribution.py:7-33).
The `torch.rand` draw keeps the reference's shape and order so that seeded runs consume the same stream."""
from __future__ import annotations

import torch
from torch import Tensor


def sample_discrete_distribution(pdf: Tensor, num_samples: int, eps: float = torch.finfo(torch.float32).eps):
    *batch, bucket = pdf.shape
    normalized_pdf = pdf / (eps + pdf.sum(dim=-1, keepdim=True))
    cdf = normalized_pdf.cumsum(dim=-1)
    samples = torch.rand((*batch, num_samples), device=pdf.device)
    index = torch.searchsorted(cdf, samples, right=True).clip(max=bucket - 1)
    return index, normalized_pdf.gather(dim=-1, index=index)


def gather_discrete_topk(pdf: Tensor, num_samples: int, eps: float = torch.finfo(torch.float32).eps):
    normalized_pdf = pdf / (eps + pdf.sum(dim=-1, keepdim=True))
    index = pdf.topk(k=num_samples, dim=-1).indices
    return index, normalized_pdf.gather(dim=-1, index=index)
