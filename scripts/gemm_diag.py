import os, sys, torch
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, R)
from latentsplat_b200.gemm import gemm_tf32
dev = torch.device('cuda:0')
torch.set_printoptions(linewidth=200, precision=2, sci_mode=False)
def run(M, N, K, a_mn=False, b_mn=False, A=None, B=None):
    A = torch.randn((K, M) if a_mn else (M, K), device=dev) if A is None else A
    B = torch.randn((K, N) if b_mn else (N, K), device=dev) if B is None else B
    y = gemm_tf32(A, B, M=M, N=N, K=K, a_mn=a_mn, b_mn=b_mn, split_k=1)
    torch.cuda.synchronize()
    ref = ((A.T if a_mn else A).double() @ (B.T if b_mn else B).double().T)
    return y, ref
# 1. identity probe, K-major: A = [I_32 ; 0], B[n][k] = n + 100 k  -> y[m][n] = B[n][m] (m < 32)
M = N = 128; K = 32
A = torch.zeros(M, K, device=dev); A[:32, :32] = torch.eye(32, device=dev)
B = (torch.arange(N, device=dev)[:, None] + 100.0 * torch.arange(K, device=dev)[None, :]).float()
y, ref = run(M, N, K, A=A, B=B)
print("K-major identity probe: y[0:4, 0:8]\n", y[:4, :8].cpu(), "\nexpected\n", ref[:4, :8].float().cpu())
print("y[30:34, 0:4]\n", y[30:34, :4].cpu(), "\nexpected\n", ref[30:34, :4].float().cpu())
# which (m,n) entries are right?
ok = (y.double() - ref).abs() < 1e-2 * (1 + ref.abs())
print("fraction correct", ok.float().mean().item(), "rows all-correct:", ok.all(1).nonzero().flatten().tolist()[:40])
# 2. ones probe: A = ones, B = ones -> y = K
y, ref = run(128, 128, 32, A=torch.ones(128, 32, device=dev), B=torch.ones(128, 32, device=dev))
print("ones probe: unique values", torch.unique(y).cpu().tolist()[:10])
y, ref = run(128, 128, 256, A=torch.ones(128, 256, device=dev), B=torch.ones(128, 256, device=dev))
print("ones probe K=256: unique values", torch.unique(y).cpu().tolist()[:10])
# 3. random, error stats per variant
for a_mn, b_mn in ((False, False), (False, True), (True, False), (True, True)):
    y, ref = run(256, 256, 128, a_mn, b_mn)
    e = (y.double() - ref).abs()
    print(f"a_mn={a_mn} b_mn={b_mn}: max err {e.max().item():.3f}  median {e.median().item():.4f}  ref rms {ref.pow(2).mean().sqrt().item():.2f}  nan {torch.isnan(y).any().item()}")
# 4. per-k probe: A[m][k] = 1 only for column k0, B = k index -> tells which k-slices are consumed
for k0 in (0, 7, 8, 15, 16, 31):
    A = torch.zeros(128, 32, device=dev); A[:, k0] = 1
    B = torch.arange(32, device=dev).float()[None, :].repeat(128, 1) + 1
    y, ref = run(128, 128, 32, A=A, B=B)
    print(f"k0={k0}: y[0,0]={y[0,0].item():.1f} expected {ref[0,0].item():.1f}; y[5,9]={y[5,9].item():.1f}")
