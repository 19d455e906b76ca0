"""Data-parallel plumbing of the hot path: scene pairs shard over ranks (one process per GPU, `torchrun`), weights
are replicated, and the only exchange is ONE all-reduce of a flat fp32 gradient buffer per optimiser step
(the reference gets the same semantics from Lightning's DDP wrapper, /root/reference/src/main.py:98, with
`find_unused_parameters=True` because loss groups switch sub-modules on and off; here the ACTIVE parameter set is
explicit and every gradient lives in one buffer, zero-filled at the start of the step, so unused parameters simply
contribute zeros -- no graph traversal, CUDA-graph friendly).

NCCL over NVLink/NVSwitch on GPUs; the same code runs on gloo for the CPU tests.
"""
from __future__ import annotations

from typing import Iterable, List

import torch
import torch.distributed as dist
from torch import Tensor, nn


class FlatGradients:
    """One contiguous gradient buffer whose slices are installed as `.grad` of the given parameters.

    autograd accumulates into pre-existing `.grad` tensors in place, so after `zero()` + backward the buffer holds
    the step's gradients and `all_reduce_mean()` is a single collective over `numel` floats."""

    def __init__(self, params: Iterable[nn.Parameter]):
        self.params: List[nn.Parameter] = [p for p in params if p.requires_grad]
        if not self.params:
            raise ValueError("no trainable parameters")
        dev, dtype = self.params[0].device, self.params[0].dtype
        self.numel = sum(p.numel() for p in self.params)
        self.flat = torch.zeros(self.numel, device=dev, dtype=dtype)
        off = 0
        for p in self.params:
            # same strides as the parameter (dense layouts only, e.g. contiguous or channels_last) -- fused optimizers
            # require param / grad layouts to match
            seg = self.flat[off:off + p.numel()]
            dense = p.is_contiguous() or (p.dim() == 4 and p.is_contiguous(memory_format=torch.channels_last))
            if not dense:
                raise ValueError("FlatGradients needs dense (contiguous or channels_last) parameters")
            p.grad = seg.as_strided(p.shape, p.stride())
            off += p.numel()

    def zero(self) -> None:
        self.flat.zero_()

    def all_reduce_mean(self, group=None) -> None:
        """SUM over ranks, then / world (DDP's gradient averaging)."""
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=group)
            self.flat.div_(dist.get_world_size(group))

    def norm(self) -> Tensor:
        return self.flat.norm()

    def clip_(self, max_norm: float) -> Tensor:
        """Gradient-norm clipping on the flat buffer (`clip_gradients(0.5, norm)`, model_wrapper.py:442-446)."""
        total = self.flat.norm()
        self.flat.mul_(torch.clamp(max_norm / (total + 1e-6), max=1.0))
        return total


def shard_batch(batch: dict, rank: int, world: int) -> dict:
    """Rows rank::world of every tensor's leading (scene) dimension -- scene pairs never interact."""
    return {k: (shard_batch(v, rank, world) if isinstance(v, dict) else v[rank::world]) for k, v in batch.items()}
