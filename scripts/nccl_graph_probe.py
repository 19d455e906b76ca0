#!/usr/bin/env python
"""2-GPU probe: BucketedAllReduce with NCCL (a) eagerly, (b) captured into a CUDA graph.  Prints a line per stage (flushed), so a
hang shows where.  torchrun --nproc-per-node 2 scripts/nccl_graph_probe.py [global|thread_local|relaxed]"""
import os
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch
import torch.distributed as dist
from torch import nn

from latentsplat_b200.parallel import BucketedAllReduce, FlatGradients


def say(*a):
    print(f"[rank {dist.get_rank()}]", *a, flush=True)


def main():
    mode = sys.argv[1] if len(sys.argv) > 1 else "thread_local"
    rank = int(os.environ["RANK"])
    torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
    dist.init_process_group("nccl")
    dev = torch.device("cuda")
    torch.manual_seed(0)
    model = nn.Sequential(*[m for _ in range(6) for m in (nn.Linear(1024, 1024), nn.GELU())]).to(dev)
    fg = FlatGradients(model.parameters())
    red = BucketedAllReduce(fg, bucket_bytes=4 << 20)
    x = torch.randn(64, 1024, device=dev) + rank

    def step():
        fg.zero()
        loss = model(x).square().mean()
        red.begin()
        loss.backward()
        red.finish()
        return loss.detach()

    say("buckets", len(red.buckets))
    for i in range(3):
        step()
    torch.cuda.synchronize()
    ref = fg.flat.clone()
    say("eager ok, grad norm", float(ref.norm()))
    t = ref.clone()
    dist.all_reduce(t)
    say("plain all_reduce ok")

    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        step()
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    say("side-stream step ok; capturing with capture_error_mode =", mode)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, capture_error_mode=mode):
        out = step()
    say("captured")
    for i in range(3):
        g.replay()
    torch.cuda.synchronize()
    say("replayed; max diff vs eager", float((fg.flat - ref).abs().max()), "loss", float(out))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
