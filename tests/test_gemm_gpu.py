"""GPU: the tcgen05 TF32 GEMM (through the C ABI) against a float64 torch reference.
Tolerance: the tensor core TRUNCATES fp32 operands to TF32 (10 mantissa bits: relative error up to 2^-10 per
operand, e.g. 3001 -> 3000), so a product is off by up to 2^-9 and a length-K dot product of unit-variance
terms by ~ sqrt(K) * 2e-3 * (a small constant); the tests allow 6e-3 * sqrt(K) * rms(A) * rms(B)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _ref(A, B, a_mn, b_mn, bias, act):
    a = (A.T if a_mn else A).double()
    b = (B.T if b_mn else B).double()
    y = a @ b.T
    if bias is not None:
        y = y + bias.double()
    if act == "relu":
        y = y.relu()
    if act == "gelu":
        y = torch.nn.functional.gelu(y)
    return y


@pytest.mark.parametrize("M,N,K", [(128, 128, 32), (128, 128, 256), (256, 384, 128), (200, 136, 100), (8200, 2304, 768),
                                   (1, 128, 20), (333, 64, 40), (4096, 156, 128)])
@pytest.mark.parametrize("a_mn,b_mn", [(False, False), (False, True), (True, True), (True, False)])
def test_gemm_layouts_and_tails(cuda, M, N, K, a_mn, b_mn):
    from latentsplat_b200.gemm import gemm_tf32
    if (a_mn and M % 4) or (b_mn and N % 4) or (not a_mn and K % 4) or (not b_mn and K % 4):
        pytest.skip("TMA needs leading dimensions that are multiples of 4 elements")
    g = torch.Generator(cuda).manual_seed(M * 7 + N * 3 + K)
    A = torch.randn((K, M) if a_mn else (M, K), device=cuda, generator=g)
    B = torch.randn((K, N) if b_mn else (N, K), device=cuda, generator=g)
    y = gemm_tf32(A, B, M=M, N=N, K=K, a_mn=a_mn, b_mn=b_mn, split_k=1)
    torch.cuda.synchronize()
    ref = _ref(A, B, a_mn, b_mn, None, "none")
    err = (y.double() - ref).abs().max().item()
    assert err <= 6e-3 * np.sqrt(K) + 1e-5, f"max err {err:.3e} (K={K})"


@pytest.mark.parametrize("act", ["none", "relu", "gelu"])
def test_gemm_fused_bias_activation(cuda, act):
    from latentsplat_b200.gemm import gemm_tf32
    g = torch.Generator(cuda).manual_seed(5)
    M, N, K = 1000, 256, 128
    A, B, bias = torch.randn(M, K, device=cuda, generator=g), torch.randn(N, K, device=cuda, generator=g) * 0.1, \
        torch.randn(N, device=cuda, generator=g)
    y = gemm_tf32(A, B, M=M, N=N, K=K, bias=bias, act=act)
    ref = _ref(A, B, False, False, bias, act)
    assert (y.double() - ref).abs().max().item() <= 6e-3 * np.sqrt(128) * 0.1 + 1e-3


def test_gemm_split_k_and_accumulate(cuda):
    from latentsplat_b200.gemm import gemm_tf32
    g = torch.Generator(cuda).manual_seed(6)
    M, N, K = 128, 256, 65536      # a weight-gradient shape: few tiles, long reduction
    A, B = torch.randn(K, M, device=cuda, generator=g), torch.randn(K, N, device=cuda, generator=g)
    ref = _ref(A, B, True, True, None, "none")
    for split in (1, 7, 0):
        y = gemm_tf32(A, B, M=M, N=N, K=K, a_mn=True, b_mn=True, split_k=split)
        assert (y.double() - ref).abs().max().item() <= 6e-3 * np.sqrt(K)
    acc = torch.ones(M, N, device=cuda)
    gemm_tf32(A, B, M=M, N=N, K=K, a_mn=True, b_mn=True, split_k=4, out=acc, accumulate=True)
    assert (acc.double() - 1 - ref).abs().max().item() <= 6e-3 * np.sqrt(K)
    with pytest.raises(RuntimeError, match="activation"):
        gemm_tf32(A, B, M=M, N=N, K=K, a_mn=True, b_mn=True, split_k=4, act="relu")


def test_linear_layer_forward_backward_matches_torch(cuda):
    from latentsplat_b200 import gemm
    torch.manual_seed(0)
    lin = gemm.Linear(128, 1024, bias=True).to(cuda)
    ref = torch.nn.Linear(128, 1024, bias=True).to(cuda)
    ref.load_state_dict(lin.state_dict())
    x = torch.randn(4, 777, 128, device=cuda, requires_grad=True)
    xr = x.detach().clone().double().requires_grad_(True)
    refd = ref.double()
    y = lin(x)
    yr = refd(xr)
    w = torch.randn_like(y)
    (y * w).sum().backward()
    (yr * w.double()).sum().backward()
    tol = lambda t, k: 2e-3 * np.sqrt(k) * float(t.abs().max())
    assert (y.double() - yr).abs().max().item() <= 6e-3 * float(yr.detach().abs().max())
    assert (x.grad.double() - xr.grad).abs().max().item() <= 6e-3 * float(xr.grad.abs().max())
    assert (lin.weight.grad.double() - refd.weight.grad).abs().max().item() <= 6e-3 * float(refd.weight.grad.abs().max())
    assert (lin.bias.grad.double() - refd.bias.grad).abs().max().item() <= 1e-4 * float(refd.bias.grad.abs().max())
    # fused ReLU path
    y2 = gemm.linear(x, lin.weight, lin.bias, act="relu")
    assert (y2.double() - yr.relu()).abs().max().item() <= 6e-3 * float(yr.detach().abs().max())


def test_fused_single_query_attention_matches_reference_math(cuda):
    """ls_sq_attention_* == the explicit softmax(q k^T d^-1/2) v of attention.py:64-68, forward and backward."""
    from latentsplat_b200.attention import single_query_attention
    R, H, S, D = 777, 4, 32, 128
    g = torch.Generator(cuda).manual_seed(3)
    q = torch.randn(R, H * D, device=cuda, generator=g, requires_grad=True)
    kv = torch.randn(R, S, 2 * H * D, device=cuda, generator=g, requires_grad=True)
    w = torch.randn(R, H * D, device=cuda, generator=g)
    out = single_query_attention(q, kv, H, D ** -0.5)
    (out * w).sum().backward()
    gq, gkv = q.grad.clone(), kv.grad.clone()
    q.grad = kv.grad = None
    qd, kvd = q.double(), kv.double()
    k, v = kvd.chunk(2, dim=-1)
    split = lambda t: t.unflatten(-1, (H, D)).transpose(1, 2)            # (R, H, n, D)
    dots = torch.matmul(split(qd[:, None]), split(k).transpose(-1, -2)) * D ** -0.5
    ref = torch.matmul(dots.softmax(dim=-1), split(v)).transpose(1, 2).flatten(-2)[:, 0]
    (ref * w.double()).sum().backward()
    assert (out.double() - ref).abs().max().item() < 1e-5
    assert (gq.double() - q.grad.double()).abs().max().item() < 1e-5
    assert (gkv.double() - kv.grad.double()).abs().max().item() < 1e-5
    # ragged S (< 32)
    kv2 = torch.randn(50, 7, 2 * H * D, device=cuda, generator=g)
    q2 = torch.randn(50, H * D, device=cuda, generator=g)
    o2 = single_query_attention(q2, kv2, H, 0.1)
    k2, v2 = kv2.double().chunk(2, dim=-1)
    d2 = torch.matmul(split(q2.double()[:, None]), split(k2).transpose(-1, -2)) * 0.1
    r2 = torch.matmul(d2.softmax(dim=-1), split(v2)).transpose(1, 2).flatten(-2)[:, 0]
    assert (o2.double() - r2).abs().max().item() < 1e-5


def test_absorbed_cross_attention_equals_explicit_attention(cuda):
    """Folding to_kv into the query / output side is exact algebra: compare with softmax(q (W_k z)^T) (W_v z) in float64,
    forward and all three gradients (q, z, W_kv)."""
    from latentsplat_b200.attention import absorbed_cross_attention
    R, H, S, D = 1500, 4, 32, 128
    g = torch.Generator(cuda).manual_seed(4)
    q = torch.randn(R, H * D, device=cuda, generator=g, requires_grad=True)
    z = torch.randn(R, S, 128, device=cuda, generator=g, requires_grad=True)
    w = (torch.randn(2 * H * D, 128, device=cuda, generator=g) / 128 ** 0.5).requires_grad_(True)
    wt = torch.randn(R, H * D, device=cuda, generator=g)
    out = absorbed_cross_attention(q, z, w, H, D ** -0.5)
    (out * wt).sum().backward()
    got = [out.detach().double(), q.grad.double(), z.grad.double(), w.grad.double()]
    q.grad = z.grad = w.grad = None
    qd, zd, wd = q.double(), z.double(), w.double()
    k, v = (zd @ wd.T).chunk(2, dim=-1)
    split = lambda t: t.unflatten(-1, (H, D)).transpose(1, 2)
    dots = torch.matmul(split(qd[:, None]), split(k).transpose(-1, -2)) * D ** -0.5
    ref = torch.matmul(dots.softmax(dim=-1), split(v)).transpose(1, 2).flatten(-2)[:, 0]
    (ref * wt.double()).sum().backward()
    want = [ref.detach(), q.grad.double(), z.grad.double(), w.grad.double()]
    for a, b, name in zip(got, want, ("out", "dq", "dz", "dW_kv")):
        err = (a - b).abs().max().item()
        assert err <= 2e-2 * b.abs().max().item(), f"{name}: {err:.3e} vs {b.abs().max().item():.3e}"     # TF32 GEMMs
        assert (a - b).abs().mean().item() <= 2e-3 * b.abs().mean().item() + 1e-6, name
