import os, sys, torch
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, R)
from latentsplat_b200.gemm import gemm_tf32
dev = torch.device('cuda:0')
M, N, K = [int(x) for x in sys.argv[1:4]]
A = torch.randn(M, K, device=dev); B = torch.randn(N, K, device=dev); out = torch.empty(M, N, device=dev)
for _ in range(4): gemm_tf32(A, B, M=M, N=N, K=K, out=out)
torch.cuda.synchronize()
