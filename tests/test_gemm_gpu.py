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


@pytest.mark.parametrize("with_encoding", [True, False])
def test_epipolar_gather_matches_grid_sample_path(cuda, with_encoding):
    """ls_epipolar_gather_* == index -> F.grid_sample(bilinear, zeros, align_corners=False) -> mask (+ Linear(PE(depth)))
    (epipolar_sampler.py:96-112, epipolar_transformer.py:121-122), forward and the three gradients."""
    import torch.nn.functional as F
    from latentsplat_b200.epipolar_gather import epipolar_gather
    from latentsplat_b200.model.encodings.positional_encoding import PositionalEncoding
    g = torch.Generator(cuda).manual_seed(11)
    images, H, W, rows, S = 6, 16, 12, 700, 32
    feat = torch.randn(images, H, W, 128, device=cuda, generator=g, requires_grad=True)
    xy = torch.rand(rows, S, 2, device=cuda, generator=g) * 1.4 - 0.2           # some samples fall outside the image
    xy[5, 3] = 0.0
    depth = torch.rand(rows, S, device=cuda, generator=g)
    image = torch.randint(0, images, (rows,), device=cuda, generator=g).to(torch.int32)
    valid = (torch.rand(rows, device=cuda, generator=g) > 0.2).float()
    pe = PositionalEncoding(10).to(cuda)
    weight = (torch.randn(128, 20, device=cuda, generator=g) * 0.3).requires_grad_(True) if with_encoding else None
    bias = torch.randn(128, device=cuda, generator=g).requires_grad_(True) if with_encoding else None
    wt = torch.randn(rows, S, 128, device=cuda, generator=g)

    z = epipolar_gather(feat, xy, depth if with_encoding else None, image, valid, weight, bias)
    (z * wt).sum().backward()
    got = [z.detach(), feat.grad.clone()] + ([weight.grad.clone(), bias.grad.clone()] if with_encoding else [])
    feat.grad = None
    if with_encoding:
        weight.grad = bias.grad = None

    nchw = feat.permute(0, 3, 1, 2)[image.long()]                                # (rows, C, H, W)
    ref = F.grid_sample(nchw, (2 * xy - 1)[:, :, None], mode="bilinear", padding_mode="zeros", align_corners=False)
    ref = ref[..., 0].transpose(1, 2) * valid[:, None, None]
    if with_encoding:
        ref = ref + F.linear(pe(depth[..., None]), weight, bias)
    (ref * wt).sum().backward()
    want = [ref.detach(), feat.grad] + ([weight.grad, bias.grad] if with_encoding else [])
    for a, b, name in zip(got, want, ("z", "dfeat", "dW", "db")):
        scale = b.abs().max().item()
        assert (a - b).abs().max().item() <= 2e-4 * scale + 1e-5, f"{name}: {(a - b).abs().max().item():.3e} of {scale:.3e}"


def test_epipolar_transformer_fused_gather_equals_explicit_path(cuda):
    """The whole module with the fused gather == the explicit PyTorch sequence (same weights, same inputs)."""
    from latentsplat_b200 import epipolar_gather as eg
    from latentsplat_b200.configs import build_modules
    from helpers import camera
    torch.manual_seed(0)
    _, enc, _, _ = build_modules(with_discriminator=True)
    et = enc.epipolar_transformer.to(cuda)
    b, v, h, w = 1, 2, 32, 32
    feats = torch.randn(b, v, 128, h, w, device=cuda)
    ex = torch.eye(4, device=cuda).repeat(b, v, 1, 1)
    ex[:, 1, 0, 3] = 0.3
    ex[:, 1, 2, 3] = 0.05
    intr = torch.tensor([[0.9, 0, 0.5], [0, 0.9, 0.5], [0, 0, 1.0]], device=cuda).repeat(b, v, 1, 1)
    near, far = torch.full((b, v), 0.5, device=cuda), torch.full((b, v), 50.0, device=cuda)
    outs = []
    for flag in (True, False):
        eg.ENABLED = flag
        try:
            f = feats.clone().requires_grad_(True)
            y, _ = et(f, ex, intr, near, far)
            y.square().mean().backward()
            outs.append((y.detach(), f.grad.clone(), et.depth_encoding[1].weight.grad.clone()))
            et.zero_grad(set_to_none=True)
        finally:
            eg.ENABLED = True
    for a, b_, name in zip(outs[0], outs[1], ("features", "d input", "d depth_encoding.weight")):
        err, scale = (a - b_).abs().max().item(), b_.abs().max().item()
        assert err <= 2e-2 * scale, f"{name}: {err:.3e} of {scale:.3e}"              # TF32 GEMMs on both sides
        assert (a - b_).abs().mean().item() <= 3e-3 * b_.abs().mean().item() + 1e-7, name


def test_dino_bf16_attention_core_matches_fp32_path(cuda):
    """_Bf16AttentionCore (one packed bf16 copy -> library flash kernel -> packed gradient) against the plain fp32
    SDPA route of the same block, forward and gradients."""
    from latentsplat_b200.model.encoder.backbone import dino_vit
    torch.manual_seed(0)
    attn = dino_vit.Attention(384, 6).to(cuda)
    x = torch.randn(3, 197, 384, device=cuda)
    res = []
    for flag in (True, False):
        dino_vit.ATTENTION_BF16 = flag
        try:
            xi = x.clone().requires_grad_(True)
            y = attn(xi)
            (y * torch.linspace(-1, 1, 384, device=cuda)).sum().backward()
            res.append((y.detach(), xi.grad.clone(), attn.qkv.weight.grad.clone()))
            attn.zero_grad(set_to_none=True)
        finally:
            dino_vit.ATTENTION_BF16 = True
    with torch.no_grad():
        assert torch.isfinite(attn(x)).all()                                     # no-grad path builds no inner graph
    for a, b, name in zip(res[0], res[1], ("out", "dx", "dWqkv")):
        err, scale = (a - b).abs().max().item(), b.abs().max().item()
        assert err <= 3e-2 * scale, f"{name}: {err:.3e} of {scale:.3e}"                # bf16 q/k/v/p


@pytest.mark.parametrize("shape,act", [((4, 128, 64, 64), "silu"), ((2, 512, 16, 16), "silu"), ((3, 64, 40, 36), "none"),
                                       ((2, 256, 1024), "none")])
def test_groupnorm_silu_matches_torch(cuda, shape, act):
    """ls_groupnorm_* == F.silu(F.group_norm(x, 32, w, b, 1e-6)) forward and dx / dgamma / dbeta (float64 reference)."""
    import torch.nn.functional as F
    from latentsplat_b200.norm import group_norm
    g = torch.Generator(cuda).manual_seed(7)
    x = (torch.randn(shape, device=cuda, generator=g) * 2 + 3).requires_grad_(True)        # non-zero mean: tests the fp64 combine
    w = (torch.rand(shape[1], device=cuda, generator=g) + 0.5).requires_grad_(True)
    b = torch.randn(shape[1], device=cuda, generator=g).requires_grad_(True)
    wt = torch.randn(shape, device=cuda, generator=g)
    y = group_norm(x, 32, w, b, 1e-6, act)
    (y * wt).sum().backward()
    got = [y.detach().double(), x.grad.double(), w.grad.double(), b.grad.double()]
    x.grad = w.grad = b.grad = None
    ref = F.group_norm(x.double(), 32, w.double(), b.double(), 1e-6)
    ref = F.silu(ref) if act == "silu" else ref
    (ref * wt.double()).sum().backward()
    want = [ref.detach(), x.grad.double(), w.grad.double(), b.grad.double()]
    for a, r, name in zip(got, want, ("y", "dx", "dgamma", "dbeta")):
        assert (a - r).abs().max().item() <= 2e-5 * r.abs().max().item() + 1e-6, f"{name}: {(a - r).abs().max().item():.3e}"


@pytest.mark.parametrize("shape", [(8, 1025, 768), (4099, 1, 128), (37, 384)])
def test_layernorm_matches_torch(cuda, shape):
    """ls_layernorm_* == F.layer_norm (float64 reference), forward and dx / dgamma / dbeta."""
    import torch.nn.functional as F
    from latentsplat_b200.norm import layer_norm
    g = torch.Generator(cuda).manual_seed(9)
    x = (torch.randn(shape, device=cuda, generator=g) * 1.5 + 2).requires_grad_(True)
    w = (torch.rand(shape[-1], device=cuda, generator=g) + 0.5).requires_grad_(True)
    b = torch.randn(shape[-1], device=cuda, generator=g).requires_grad_(True)
    wt = torch.randn(shape, device=cuda, generator=g)
    y = layer_norm(x, w, b, 1e-6)
    (y * wt).sum().backward()
    got = [y.detach().double(), x.grad.double(), w.grad.double(), b.grad.double()]
    x.grad = w.grad = b.grad = None
    ref = F.layer_norm(x.double(), shape[-1:], w.double(), b.double(), 1e-6)
    (ref * wt.double()).sum().backward()
    want = [ref.detach(), x.grad.double(), w.grad.double(), b.grad.double()]
    for a, r, name in zip(got, want, ("y", "dx", "dgamma", "dbeta")):
        assert (a - r).abs().max().item() <= 2e-5 * r.abs().max().item() + 1e-6, f"{name}: {(a - r).abs().max().item():.3e}"


def test_grouped_linear_matches_baddbmm(cuda):
    """gemm.grouped_linear (one autograd node, G GEMMs into slices) == baddbmm in float64: y, dx, dW, db."""
    from latentsplat_b200.gemm import grouped_linear
    g = torch.Generator(cuda).manual_seed(2)
    x = torch.randn(3, 1000, 128, device=cuda, generator=g, requires_grad=True)
    w = (torch.randn(3, 156, 128, device=cuda, generator=g) / 128 ** 0.5).requires_grad_(True)
    b = torch.randn(3, 156, device=cuda, generator=g, requires_grad=True)
    wt = torch.randn(3, 1000, 156, device=cuda, generator=g)
    y = grouped_linear(x, w, b)
    (y * wt).sum().backward()
    got = [y.detach().double(), x.grad.double(), w.grad.double(), b.grad.double()]
    x.grad = w.grad = b.grad = None
    ref = torch.baddbmm(b.double()[:, None], x.double(), w.double().transpose(1, 2))
    (ref * wt.double()).sum().backward()
    want = [ref.detach(), x.grad.double(), w.grad.double(), b.grad.double()]
    for a, r, name in zip(got, want, ("y", "dx", "dW", "db")):
        assert (a - r).abs().max().item() <= 2e-2 * r.abs().max().item(), f"{name}: {(a - r).abs().max().item():.3e}"   # TF32


@pytest.mark.parametrize("rows,cols,ld", [(8200, 768, 768), (65536, 156, 156), (1000, 158, 160), (7, 4, 4)])
def test_col_sum_matches_torch(cuda, rows, cols, ld):
    from latentsplat_b200.gemm import col_sum
    g = torch.Generator(cuda).manual_seed(1)
    x = torch.randn(rows, ld, device=cuda, generator=g)[:, :cols]
    got, want = col_sum(x).double(), x.double().sum(dim=0)
    assert got.shape == want.shape
    assert (got - want).abs().max().item() <= 1e-5 * x.abs().double().sum(dim=0).max().item() + 1e-6


@pytest.mark.parametrize("act", ["none", "gelu", "silu"])
def test_linear_fused_activation_and_residual(cuda, act):
    """gemm.linear(act=, residual=): act(x W^T + b) + residual in ONE GEMM epilogue (+ the pre-activation kept for backward);
    output and all four gradients against float64 torch."""
    from latentsplat_b200.gemm import linear
    g = torch.Generator(cuda).manual_seed(8)
    x = torch.randn(3, 333, 128, device=cuda, generator=g, requires_grad=True)
    w = (torch.randn(256, 128, device=cuda, generator=g) / 128 ** 0.5).requires_grad_(True)
    b = torch.randn(256, device=cuda, generator=g, requires_grad=True)
    r = torch.randn(3, 333, 256, device=cuda, generator=g, requires_grad=True)
    wt = torch.randn(3, 333, 256, device=cuda, generator=g)
    y = linear(x, w, b, act=act, residual=r)
    gx, gw, gb, gr = torch.autograd.grad(y, (x, w, b, r), wt)
    xd, wd, bd, rd = (t.detach().double().requires_grad_(True) for t in (x, w, b, r))
    fn = {"none": lambda t: t, "gelu": torch.nn.functional.gelu, "silu": torch.nn.functional.silu}[act]
    ref = fn(torch.nn.functional.linear(xd, wd, bd)) + rd
    rx, rw, rb, rr = torch.autograd.grad(ref, (xd, wd, bd, rd), wt.double())
    for a, e, name, tol in ((y, ref.detach(), "y", 2e-2), (gx, rx, "dx", 2e-2), (gw, rw, "dW", 2e-2), (gb, rb, "db", 1e-3), (gr, rr, "dres", 1e-6)):
        assert (a.double() - e).abs().max().item() <= tol * e.abs().max().item() + 1e-6, f"{name}: {(a.double() - e).abs().max().item():.3e}"
