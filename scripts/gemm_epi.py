"""Cost of the fused GEMM epilogues on the DINO shapes: none vs gelu(+pre_out) vs residual."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch
from latentsplat_b200.gemm import gemm_tf32

dev = torch.device("cuda:0")
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timed(fn, reps=7):
    fn(); fn()
    ts = []
    for _ in range(reps):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    return sorted(ts)[len(ts) // 2]


for M, N, K in [(8200, 3072, 768), (8200, 768, 3072), (8200, 768, 768), (8200, 2304, 768)]:
    A, B = torch.randn(M, K, device=dev), torch.randn(N, K, device=dev) * 0.03
    bias, res = torch.randn(N, device=dev), torch.randn(M, N, device=dev)
    out, pre = torch.empty(M, N, device=dev), torch.empty(M, N, device=dev)
    r = {
        "bias": timed(lambda: gemm_tf32(A, B, M=M, N=N, K=K, bias=bias, out=out)),
        "gelu": timed(lambda: gemm_tf32(A, B, M=M, N=N, K=K, bias=bias, act="gelu", out=out)),
        "gelu+pre": timed(lambda: gemm_tf32(A, B, M=M, N=N, K=K, bias=bias, act="gelu", out=out, pre_out=pre)),
        "residual": timed(lambda: gemm_tf32(A, B, M=M, N=N, K=K, bias=bias, out=out, residual=res)),
    }
    fl = 2.0 * M * N * K
    print(M, N, K, {k: (round(v * 1e3, 1), round(fl / v / 1e9)) for k, v in r.items()}, flush=True)
