"""Sampling of depth buckets (/root/reference/src/misc/discrete_probability_distribution.py:7-33).
The `torch.rand` draw keeps the reference's shape and order so that seeded runs consume the same stream."""
from __future__ import annotations

import torch
from torch import Tensor


def sample_discrete_distribution(pdf: Tensor, num_samples: int, eps: float = torch.finfo(torch.float32).eps):
    *batch, bucket = pdf.shape
    normalized_pdf = pdf / (eps + pdf.sum(dim=-1, keepdim=True))
    with torch.no_grad():       # the CDF only feeds searchsorted (no gradient path)
        if pdf.is_cuda and bucket <= 64:
            # torch's innermost-dim scan spends 2.8 ms on 524 288 rows x 32 buckets; the same prefix sums as one fp32
            # (SIMT, not TF32) matmul with an upper-triangular ones matrix take < 0.1 ms
            tri = torch.ones(bucket, bucket, device=pdf.device, dtype=pdf.dtype).triu_()
            cdf = (normalized_pdf.reshape(-1, bucket) @ tri).reshape(normalized_pdf.shape)
        else:
            cdf = normalized_pdf.cumsum(dim=-1)
    samples = torch.rand((*batch, num_samples), device=pdf.device)
    index = torch.searchsorted(cdf, samples, right=True).clip(max=bucket - 1)
    return index, normalized_pdf.gather(dim=-1, index=index)


def gather_discrete_topk(pdf: Tensor, num_samples: int, eps: float = torch.finfo(torch.float32).eps):
    normalized_pdf = pdf / (eps + pdf.sum(dim=-1, keepdim=True))
    index = pdf.topk(k=num_samples, dim=-1).indices
    return index, normalized_pdf.gather(dim=-1, index=index)
