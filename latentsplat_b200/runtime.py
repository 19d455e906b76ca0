"""CUDA-graph execution of a static-shape step (B200-first: graphs instead of a tracing compiler).

The render path launches ~140 small kernels per step (camera set-up, the 7 rasterizer kernels, loss heads,
autograd glue); eager PyTorch spends ~5x the GPU time of the step in Python/launch latency
(profiles/r01_*).  `GraphedStep` captures forward + backward once and replays it with one launch.
Requirements on the captured function: static shapes, no host syncs (use the rasterizer's sync-free
`capacity` mode), no host-side randomness.
"""
from __future__ import annotations

from typing import Callable, Dict

import torch
from torch import Tensor

from .rasterizer import check_overflow, reserve_overflow_slots


class GraphedStep:
    """`fn(inputs: dict[str, Tensor]) -> dict[str, Tensor]` captured into one CUDA graph.

    `example_inputs` fixes shapes/dtypes; their storage becomes the graph's static input buffers.
    Calling the object copies new inputs (device or pinned-host tensors, non-blocking) into those buffers,
    replays the graph and returns the static output tensors (valid until the next call)."""

    def __init__(self, fn: Callable[[Dict[str, Tensor]], Dict[str, Tensor]], example_inputs: Dict[str, Tensor],
                 warmup: int = 3):
        self.fn = fn
        self.static_in = {k: v.clone() for k, v in example_inputs.items()}
        reserve_overflow_slots()              # pinned flag buffers for the rasterizer calls about to be captured
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(warmup):
                self.fn(self.static_in)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.static_out = self.fn(self.static_in)

    def load(self, inputs: Dict[str, Tensor]) -> None:
        for k, v in inputs.items():
            self.static_in[k].copy_(v, non_blocking=True)

    def replay(self) -> Dict[str, Tensor]:
        # a captured sync-free rasterizer call copies its overflow flag to pinned memory inside the graph: an overflow of an
        # earlier replay raises here (rasterizer._OverflowGuard), one step late at most
        check_overflow()
        self.graph.replay()
        return self.static_out

    def __call__(self, inputs: Dict[str, Tensor] | None = None) -> Dict[str, Tensor]:
        if inputs:
            self.load(inputs)
        return self.replay()
