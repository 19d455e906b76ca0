"""Compute-bound GEMM (4096^3) in the four operand layouts: is the MN-major (32-B atom) path slower in the tensor core?"""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch
from latentsplat_b200.gemm import gemm_tf32

dev = torch.device("cuda:0")
M = N = K = 4096
for a_mn in (False, True):
    for b_mn in (False, True):
        A = torch.randn((K, M) if a_mn else (M, K), device=dev)
        B = torch.randn((K, N) if b_mn else (N, K), device=dev)
        out = torch.empty(M, N, device=dev)
        for _ in range(3):
            gemm_tf32(A, B, M=M, N=N, K=K, a_mn=a_mn, b_mn=b_mn, out=out, split_k=1)
        ts = []
        for _ in range(7):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); gemm_tf32(A, B, M=M, N=N, K=K, a_mn=a_mn, b_mn=b_mn, out=out, split_k=1); b.record()
            torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
        t = sorted(ts)[3]
        print(f"a_mn={a_mn} b_mn={b_mn}: {t:.3f} ms  {2*M*N*K/t/1e9:.0f} TF/s", flush=True)
