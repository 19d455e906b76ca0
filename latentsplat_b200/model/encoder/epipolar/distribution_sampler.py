"""Bucket sampling / gathering (/root/reference/src/model/encoder/epipolar/distribution_sampler.py:11-51)."""
from __future__ import annotations

import torch
from torch import Tensor

from ....misc.discrete_probability_distribution import gather_discrete_topk, sample_discrete_distribution


class DistributionSampler:
    def sample(self, pdf: Tensor, deterministic: bool, num_samples: int):
        if deterministic:
            return gather_discrete_topk(pdf, num_samples)
        return sample_discrete_distribution(pdf, num_samples)

    def gather(self, index: Tensor, target: Tensor) -> Tensor:
        """Gather along the bucket dimension (= last dimension of `index`), broadcasting trailing shape."""
        bucket_dim = index.ndim - 1
        while index.ndim < target.ndim:
            index = index[..., None]
        shape = list(target.shape)
        shape[bucket_dim] = index.shape[bucket_dim]
        index = index.broadcast_to(shape)
        if target.shape[bucket_dim] == 1:
            index = torch.zeros_like(index)
        return target.gather(dim=bucket_dim, index=index)
