"""A/B of the 1-CTA and the experimental 2-CTA GEMM on a few shapes (run twice: with and without LS_GEMM_2CTA=1)."""
import os, sys, torch
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, R)
from latentsplat_b200.gemm import gemm_tf32
dev = torch.device('cuda:0')
flush = torch.empty(64 * 1024 * 1024, device=dev)
tag = "2cta" if os.environ.get("LS_GEMM_2CTA") == "1" else "1cta"
for (M, N, K, a_mn, b_mn) in [(8200, 3072, 768, False, False), (8200, 2304, 768, False, False), (8200, 768, 3072, False, False),
                              (8200, 768, 3072, False, True), (3072, 768, 8200, True, True)]:
    A = torch.randn((K, M) if a_mn else (M, K), device=dev); B = torch.randn((K, N) if b_mn else (N, K), device=dev)
    out = torch.empty(M, N, device=dev)
    for _ in range(3): gemm_tf32(A, B, M=M, N=N, K=K, a_mn=a_mn, b_mn=b_mn, out=out, split_k=1)
    ts = []
    for _ in range(10):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); gemm_tf32(A, B, M=M, N=N, K=K, a_mn=a_mn, b_mn=b_mn, out=out, split_k=1); e1.record()
        torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    ms = sorted(ts)[len(ts) // 2]
    ref = (A.t() if a_mn else A) @ (B if b_mn else B.t())
    err = ((out - ref).abs().max() / ref.abs().max()).item()
    print(f"{tag} M={M} N={N} K={K} a_mn={int(a_mn)} b_mn={int(b_mn)}: {ms*1e3:.1f} us  {2*M*N*K/ms/1e9:.0f} TF/s  relerr {err:.1e}", flush=True)
