import os, sys, torch
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, R)
from latentsplat_b200.gemm import gemm_tf32
dev = torch.device('cuda:0')
def t(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n
shapes = [("dino qkv", 8200, 2304, 768, 0, 0), ("dino fc1", 8200, 3072, 768, 0, 0), ("dino fc2", 8200, 768, 3072, 0, 0),
          ("to_kv fwd", 1048576, 1024, 128, 0, 0), ("to_kv dgrad", 1048576, 128, 1024, 0, 1), ("to_kv wgrad", 1024, 128, 1048576, 1, 1),
          ("to_gauss", 524288, 156, 128, 0, 0), ("attn to_out", 32768, 128, 512, 0, 0), ("dino fc1 wgrad", 3072, 768, 8200, 1, 1)]
print(f"{'shape':16s} {'M':>8s} {'N':>6s} {'K':>8s} | ours ms  TF/s  GB/s | cublas-tf32 ms | cublas-fp32 ms")
for name, M, N, K, amn, bmn in shapes:
    A = torch.randn((K, M) if amn else (M, K), device=dev); B = torch.randn((K, N) if bmn else (N, K), device=dev)
    out = torch.empty(M, N, device=dev)
    ours = t(lambda: gemm_tf32(A, B, M=M, N=N, K=K, a_mn=bool(amn), b_mn=bool(bmn), out=out))
    Ar = A.T if amn else A; Br = (B.T if bmn else B)
    torch.backends.cuda.matmul.allow_tf32 = True
    cb = t(lambda: torch.matmul(Ar, Br.T, out=out))
    torch.backends.cuda.matmul.allow_tf32 = False
    cf = t(lambda: torch.matmul(Ar, Br.T, out=out), n=5)
    fl = 2.0 * M * N * K; by = 4.0 * (M * K + N * K + M * N)
    print(f"{name:16s} {M:8d} {N:6d} {K:8d} | {ours:7.3f} {fl/ours/1e9:6.1f} {by/ours/1e6:6.0f} | {cb:7.3f} ({fl/cb/1e9:6.1f} TF/s) | {cf:7.3f}")
