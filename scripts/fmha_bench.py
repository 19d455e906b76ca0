#!/usr/bin/env python
"""Times the tcgen05 flash-attention kernels (forward, backward split into delta / dq / dkv through CUPTI) on the DINO ViT-B/8
shape of the full step (8 images x 12 heads x 1025 tokens x 64) and the image self-attention shape, next to torch SDPA."""
import sys
from pathlib import Path

import torch
import torch.nn.functional as F

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from latentsplat_b200 import fmha  # noqa: E402


def timeit(fn, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    dev = torch.device("cuda")
    for B, H, L, D in [(8, 12, 1025, 64), (8, 4, 256, 128)]:
        qkv = torch.randn(B, L, 3 * H * D, device=dev, requires_grad=True)
        go = torch.randn(B, L, H * D, device=dev)
        fwd = timeit(lambda: fmha.attention_packed(qkv.detach(), H, D ** -0.5))
        out = fmha.attention_packed(qkv, H, D ** -0.5)
        bwd = timeit(lambda: torch.autograd.grad(out, qkv, go, retain_graph=True))
        flops = 4.0 * B * H * L * L * D
        print(f"B={B} H={H} L={L} D={D}: ours fwd {fwd*1e3:.1f} us ({flops/fwd/1e9:.0f} TF/s)  bwd {bwd*1e3:.1f} us ({2.5*flops/bwd/1e9:.0f} TF/s)")
        from torch.profiler import ProfilerActivity, profile
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            for _ in range(5):
                torch.autograd.grad(fmha.attention_packed(qkv, H, D ** -0.5), qkv, go)
            torch.cuda.synchronize()
        for row in prof.key_averages():
            if "fmha" in row.key:
                print(f"    {row.key[:60]:60s} {row.device_time_total / row.count:9.1f} us x{row.count}")
        q, k, v = (t.transpose(1, 2).contiguous() for t in qkv.detach().view(B, L, 3, H, D).unbind(2))
        for name, cast in (("bf16", torch.bfloat16), ("fp32", torch.float32)):
            qq, kk, vv = (t.to(cast).requires_grad_(True) for t in (q, k, v))
            f = timeit(lambda: F.scaled_dot_product_attention(qq, kk, vv))
            o = F.scaled_dot_product_attention(qq, kk, vv)
            g = torch.randn_like(o)
            b = timeit(lambda: torch.autograd.grad(o, (qq, kk, vv), g, retain_graph=True))
            print(f"    torch SDPA {name}: fwd {f*1e3:.1f} us  bwd {b*1e3:.1f} us")


if __name__ == "__main__":
    main()
