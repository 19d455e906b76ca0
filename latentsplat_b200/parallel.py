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


class BucketedAllReduce:
    """Gradient averaging overlapped with the backward pass (the job DDP's reducer does for the reference, src/main.py:98).

    The flat buffer of a `FlatGradients` is cut into contiguous buckets in parameter order.  A post-accumulate hook on every
    parameter counts its bucket down; the moment a bucket is complete its all-reduce is enqueued on a communication stream that
    first waits for the stream the gradient was produced on.  Backward reaches the modules in reverse order, so the decoder /
    discriminator buckets travel while the encoder is still differentiating and only the last bucket (the first ViT blocks) is
    exposed.  `finish()` sends whatever is left (parameters that took no part in this step keep their zero-filled slices) and
    makes the current stream wait for the communication stream.

    Everything is stream-ordered and free of host syncs, so the whole begin() .. backward .. finish() sequence can be captured
    into a CUDA graph (NCCL collectives are capturable): the replay then overlaps NVLink traffic with the remaining kernels
    with no Python in between.  The same code runs on gloo (CPU tests): there the collectives are synchronous."""

    def __init__(self, fgrads: FlatGradients, bucket_bytes: int = 32 << 20, group=None):
        self.fgrads, self.group = fgrads, group
        self.world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
        self.cuda = fgrads.flat.is_cuda
        self.comm = torch.cuda.Stream(fgrads.flat.device) if self.cuda else None
        self.avg = self.cuda                                   # NCCL averages inside the collective; gloo: SUM then divide
        self.buckets: List[List[int]] = []                     # [start, end, n_params]
        self.bucket_of = {}
        off, start, count = 0, 0, 0
        limit = max(1, bucket_bytes // fgrads.flat.element_size())
        for p in fgrads.params:
            self.bucket_of[id(p)] = len(self.buckets)
            off += p.numel()
            count += 1
            if off - start >= limit:
                self.buckets.append([start, off, count])
                start, count = off, 0
        if count:
            self.buckets.append([start, off, count])
        self.remaining = [b[2] for b in self.buckets]
        self.sent = [False] * len(self.buckets)
        self.active = False
        self.handles = [p.register_post_accumulate_grad_hook(self._hook) for p in fgrads.params] if self.world > 1 else []

    def begin(self) -> None:
        """Call before the backward pass of every step (inside the captured function when the step is a CUDA graph)."""
        self.remaining = [b[2] for b in self.buckets]
        self.sent = [False] * len(self.buckets)
        self.active = self.world > 1

    def _send(self, i: int) -> None:
        self.sent[i] = True
        seg = self.fgrads.flat[self.buckets[i][0]:self.buckets[i][1]]
        if self.cuda:
            self.comm.wait_stream(torch.cuda.current_stream(seg.device))      # the gradients of this bucket are complete
            with torch.cuda.stream(self.comm):
                dist.all_reduce(seg, op=dist.ReduceOp.AVG if self.avg else dist.ReduceOp.SUM, group=self.group)
        else:
            dist.all_reduce(seg, op=dist.ReduceOp.SUM, group=self.group)
            seg.div_(self.world)

    def _hook(self, p: nn.Parameter) -> None:
        if not self.active:
            return
        i = self.bucket_of[id(p)]
        self.remaining[i] -= 1
        if self.remaining[i] == 0 and not self.sent[i]:
            self._send(i)

    def finish(self) -> None:
        """After backward: reduce the buckets that never completed and join the communication stream."""
        if not self.active:
            return
        self.active = False
        for i in range(len(self.buckets) - 1, -1, -1):
            if not self.sent[i]:
                self._send(i)
        if self.cuda:
            torch.cuda.current_stream(self.fgrads.flat.device).wait_stream(self.comm)

    def remove(self) -> None:
        for h in self.handles:
            h.remove()
        self.handles = []


def shard_batch(batch: dict, rank: int, world: int) -> dict:
    """Rows rank::world of every tensor's leading (scene) dimension -- scene pairs never interact."""
    return {k: (shard_batch(v, rank, world) if isinstance(v, dict) else v[rank::world]) for k, v in batch.items()}
