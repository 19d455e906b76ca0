"""GPU: the tcgen05 implicit-GEMM convolutions (through the C ABI, include/ls_conv.h) against float64 torch convolutions.

Tolerance: operands are truncated to TF32 by the tensor core (relative 2^-10 per operand), products accumulate in fp32;
for a reduction of K unit-variance terms the error is ~ sqrt(K) * 2e-3 * rms(a) * rms(b).  The tests allow
6e-3 * sqrt(K) * rms(a) * rms(b) + 1e-5 (the bound of tests/test_gemm_gpu.py) with K = the reduction length of the pass
(forward: R*S*Cin, dgrad: taps*Cout, wgrad: N*OH*OW)."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

CL = torch.channels_last


def _rms(t):
    return float(t.double().square().mean().sqrt())


def _check(got, want, K, ra, rb, what):
    err = (got.double() - want).abs().max().item()
    tol = 6e-3 * np.sqrt(K) * ra * rb + 1e-5
    assert err <= tol, f"{what}: max err {err:.3e} > {tol:.3e}"


# (N, Cin, H, W, Cout, k, stride, pad): shapes of the hot path at reduced size + edge cases
CONV_CASES = [
    (2, 32, 16, 16, 64, 3, 1, 1),        # VAE 3x3
    (1, 128, 32, 32, 256, 3, 1, 1),      # BN = 256 tile
    (2, 64, 20, 24, 128, 7, 1, 3),       # encoder 7x7, ragged grid (tiles overhang)
    (2, 4, 32, 32, 32, 3, 1, 1),         # conv_in: 4 input channels (zero-filled K tail)
    (3, 32, 16, 16, 4, 3, 1, 1),         # conv_out-like: few output channels (BN = 32)
    (2, 16, 32, 32, 32, 4, 2, 1),        # PatchGAN stride 2
    (2, 32, 17, 19, 48, 4, 1, 1),        # PatchGAN stride 1, odd sizes
    (2, 32, 64, 64, 32, 4, 4, 0),        # down-scaler 4x4 stride 4
    (2, 40, 8, 8, 72, 1, 1, 0),          # 1x1 shortcut, channel tails
    (1, 8, 256, 256, 16, 3, 1, 1),       # full-width rows (BW = 128)
    (2, 4, 64, 64, 96, 8, 8, 0),         # DINO patch embedding: 8x8 stride 8 on a (padded) RGB image
    (2, 256, 16, 16, 256, 3, 1, 1),      # 256 channels both ways: the cta_group::2 kernels (forward, dgrad, wgrad) when forced
    (1, 256, 32, 40, 512, 4, 2, 1),      # strided, two channel pairs, ragged width
]


def _make(case, seed, dev):
    N, Cin, H, W, Cout, k, st, pad = case
    g = torch.Generator(dev).manual_seed(seed)
    x = torch.randn(N, Cin, H, W, device=dev, generator=g).contiguous(memory_format=CL)
    w = (torch.randn(Cout, Cin, k, k, device=dev, generator=g) / np.sqrt(Cin * k * k)).contiguous(memory_format=CL)
    b = torch.randn(Cout, device=dev, generator=g)
    return x, w, b


@pytest.mark.parametrize("case", CONV_CASES, ids=lambda c: "x".join(map(str, c)))
def test_conv2d_forward(cuda, case):
    from latentsplat_b200.conv import conv2d
    N, Cin, H, W, Cout, k, st, pad = case
    x, w, b = _make(case, 1, cuda)
    y = conv2d(x, w, b, st, pad)
    torch.cuda.synchronize()
    ref = F.conv2d(x.double(), w.double(), b.double(), st, pad)
    assert y.shape == ref.shape and y.is_contiguous(memory_format=CL)
    _check(y, ref, Cin * k * k, _rms(x), _rms(w), "forward")


@pytest.mark.parametrize("case", CONV_CASES, ids=lambda c: "x".join(map(str, c)))
def test_conv2d_backward(cuda, case):
    from latentsplat_b200.conv import conv2d
    N, Cin, H, W, Cout, k, st, pad = case
    x, w, b = _make(case, 2, cuda)
    x.requires_grad_(True); w.requires_grad_(True); b.requires_grad_(True)
    y = conv2d(x, w, b, st, pad)
    gy = torch.randn(y.shape, device=cuda, generator=torch.Generator(cuda).manual_seed(3)).contiguous(memory_format=CL)
    gx, gw, gb = torch.autograd.grad(y, (x, w, b), gy)
    torch.cuda.synchronize()
    xd, wd, bd = (t.detach().double().requires_grad_(True) for t in (x, w, b))
    rx, rw, rb = torch.autograd.grad(F.conv2d(xd, wd, bd, st, pad), (xd, wd, bd), gy.double())
    taps = (-(-k // st)) ** 2                       # filter taps that reach one residue class of a strided dgrad
    _check(gx, rx, taps * Cout, _rms(gy), _rms(w), "dgrad")
    _check(gw, rw, N * y.shape[2] * y.shape[3], _rms(gy), _rms(x), "wgrad")
    assert gw.stride() == w.stride()
    assert (gb.double() - rb).abs().max().item() <= 1e-4 * rb.abs().max().item() + 1e-4


@pytest.mark.parametrize("act", ["relu", "gelu", "silu", "lrelu"])
def test_conv2d_fused_activation_forward_backward(cuda, act):
    from latentsplat_b200.conv import conv2d
    case = (2, 32, 16, 16, 64, 3, 1, 1)
    x, w, b = _make(case, 4, cuda)
    x.requires_grad_(True)
    y = conv2d(x, w, b, 1, 1, act=act)
    gy = torch.randn_like(y)
    (gx,) = torch.autograd.grad(y, x, gy)
    fn = {"relu": F.relu, "gelu": F.gelu, "silu": F.silu, "lrelu": lambda t: F.leaky_relu(t, 0.2)}[act]
    xd = x.detach().double().requires_grad_(True)
    pre = F.conv2d(xd, w.double(), b.double(), 1, 1)
    ref = fn(pre)
    _check(y, ref.detach(), 32 * 9, _rms(x), _rms(w), f"{act} forward")
    if act in ("relu", "lrelu"):
        # the kinked activations: a TF32 rounding of a pre-activation near 0 flips its derivative, so the reference gradient
        # takes the derivative mask from OUR output (what our backward must be consistent with)
        mask = torch.where(y.detach() > 0, 1.0, 0.0 if act == "relu" else 0.2).double()
        (rx,) = torch.autograd.grad(pre, xd, gy.double() * mask)
    else:
        (rx,) = torch.autograd.grad(ref, xd, gy.double())
    _check(gx, rx, 9 * 64, _rms(gy), _rms(w), f"{act} dgrad")


@pytest.mark.parametrize("N,Cin,H,W,Cout,k", [(2, 32, 16, 16, 64, 4), (1, 128, 8, 8, 128, 4), (2, 16, 5, 7, 8, 2)])
def test_conv_transpose2d_forward_backward(cuda, N, Cin, H, W, Cout, k):
    from latentsplat_b200.conv import conv2d
    g = torch.Generator(cuda).manual_seed(5)
    x = torch.randn(N, Cin, H, W, device=cuda, generator=g).contiguous(memory_format=CL).requires_grad_(True)
    w = (torch.randn(Cin, Cout, k, k, device=cuda, generator=g) / np.sqrt(Cin)).contiguous(memory_format=CL).requires_grad_(True)
    b = torch.randn(Cout, device=cuda, generator=g).requires_grad_(True)
    y = conv2d(x, w, b, k, 0, transposed=True)
    gy = torch.randn(y.shape, device=cuda, generator=g).contiguous(memory_format=CL)
    gx, gw, gb = torch.autograd.grad(y, (x, w, b), gy)
    xd, wd, bd = (t.detach().double().requires_grad_(True) for t in (x, w, b))
    ref = F.conv_transpose2d(xd, wd, bd, k)
    rx, rw, rb = torch.autograd.grad(ref, (xd, wd, bd), gy.double())
    _check(y, ref.detach(), Cin, _rms(x), _rms(w), "transposed forward")
    _check(gx, rx, k * k * Cout, _rms(gy), _rms(w), "transposed dgrad")
    _check(gw, rw, N * H * W, _rms(gy), _rms(x), "transposed wgrad")
    assert (gb.double() - rb).abs().max().item() <= 1e-4 * rb.abs().max().item() + 1e-4


@pytest.mark.parametrize("N,C,H,W,Cout", [(2, 32, 16, 16, 64), (1, 256, 8, 12, 256), (2, 64, 7, 33, 32)])
def test_upsample2x_conv3x3_matches_interpolate_then_conv(cuda, N, C, H, W, Cout):
    """The VAE up-sampler (nearest 2x, then 3x3 conv) with the up-sampling folded into the filter: output and all gradients
    against float64 F.interpolate + F.conv2d."""
    from latentsplat_b200.conv import upsample2x_conv3x3
    g = torch.Generator(cuda).manual_seed(6)
    x = torch.randn(N, C, H, W, device=cuda, generator=g).contiguous(memory_format=CL).requires_grad_(True)
    w = (torch.randn(Cout, C, 3, 3, device=cuda, generator=g) / np.sqrt(9 * C)).contiguous(memory_format=CL).requires_grad_(True)
    b = torch.randn(Cout, device=cuda, generator=g).requires_grad_(True)
    y = upsample2x_conv3x3(x, w, b)
    gy = torch.randn(y.shape, device=cuda, generator=g).contiguous(memory_format=CL)
    gx, gw, gb = torch.autograd.grad(y, (x, w, b), gy)
    xd, wd, bd = (t.detach().double().requires_grad_(True) for t in (x, w, b))
    ref = F.conv2d(F.interpolate(xd, scale_factor=2.0, mode="nearest"), wd, bd, 1, 1)
    rx, rw, rb = torch.autograd.grad(ref, (xd, wd, bd), gy.double())
    assert y.shape == ref.shape
    _check(y, ref.detach(), 9 * C, _rms(x), _rms(w), "upconv forward")
    _check(gx, rx, 36 * Cout, _rms(gy), _rms(w), "upconv dgrad")
    _check(gw, rw, 4 * N * H * W, _rms(gy), _rms(x), "upconv wgrad")
    assert (gb.double() - rb).abs().max().item() <= 1e-4 * rb.abs().max().item() + 1e-4


def test_conv_modules_match_torch_modules(cuda):
    """Conv2d / ConvTranspose2d modules (channel padding for RGB / 7-channel / 1-channel layers, NCHW inputs)."""
    from latentsplat_b200 import conv
    torch.manual_seed(0)
    for ours, ref, shape in [
        (conv.Conv2d(3, 32, 7, 1, 3), torch.nn.Conv2d(3, 32, 7, 1, 3), (2, 3, 32, 32)),
        (conv.Conv2d(7, 64, 1), torch.nn.Conv2d(7, 64, 1), (2, 7, 16, 16)),
        (conv.Conv2d(32, 1, 4, 1, 1), torch.nn.Conv2d(32, 1, 4, 1, 1), (2, 32, 9, 9)),
        (conv.Conv2d(32, 3, 3, 1, 1, bias=False), torch.nn.Conv2d(32, 3, 3, 1, 1, bias=False), (2, 32, 16, 16)),
        (conv.ConvTranspose2d(32, 32, 4, 4), torch.nn.ConvTranspose2d(32, 32, 4, 4), (2, 32, 8, 8)),
    ]:
        ours, ref = ours.to(cuda), ref.to(cuda).double()
        ref.load_state_dict({k: v.double() for k, v in ours.state_dict().items()})
        x = torch.randn(shape, device=cuda)
        xi, xr = x.clone().requires_grad_(True), x.double().requires_grad_(True)
        y, yr = ours(xi), ref(xr)
        gy = torch.randn_like(y)
        y.backward(gy); yr.backward(gy.double())
        K = ours.weight[0].numel()
        _check(y, yr.detach(), K, _rms(x), _rms(ours.weight), "module forward")
        _check(xi.grad, xr.grad, ours.weight.numel() // x.shape[1], _rms(gy), _rms(ours.weight), "module dgrad")
        _check(ours.weight.grad, ref.weight.grad, gy[:, 0].numel(), _rms(gy), _rms(x), "module wgrad")
        if ours.bias is not None:
            assert (ours.bias.grad.double() - ref.bias.grad).abs().max().item() <= 1e-4 * ref.bias.grad.abs().max().item() + 1e-4


@pytest.mark.parametrize("shape,groups,act", [((2, 64, 16, 16), 32, "silu"), ((3, 512, 8, 8), 32, "none"), ((2, 128, 33, 7), 32, "silu"),
                                               ((1, 96, 10, 10), 8, "silu")])
def test_groupnorm_nhwc_matches_float64(cuda, shape, groups, act):
    """GroupNorm (+SiLU) on channels_last activations (ls_groupnorm_nhwc_*): y, dx, dgamma, dbeta against float64 torch."""
    from latentsplat_b200.norm import group_norm
    g = torch.Generator(cuda).manual_seed(11)
    x = (torch.randn(shape, device=cuda, generator=g) * 2 + 0.5).contiguous(memory_format=CL).requires_grad_(True)
    w = (1 + 0.2 * torch.randn(shape[1], device=cuda, generator=g)).requires_grad_(True)
    b = (0.3 * torch.randn(shape[1], device=cuda, generator=g)).requires_grad_(True)
    gy = torch.randn(shape, device=cuda, generator=g).contiguous(memory_format=CL)
    y = group_norm(x, groups, w, b, 1e-6, act)
    assert y.is_contiguous(memory_format=CL)
    gx, gw, gb = torch.autograd.grad(y, (x, w, b), gy)
    xd, wd, bd = (t.detach().double().requires_grad_(True) for t in (x, w, b))
    ref = F.group_norm(xd, groups, wd, bd, 1e-6)
    if act == "silu":
        ref = F.silu(ref)
    rx, rw, rb = torch.autograd.grad(ref, (xd, wd, bd), gy.double())
    for a, r, name in ((y, ref.detach(), "y"), (gx, rx, "dx"), (gw, rw, "dgamma"), (gb, rb, "dbeta")):
        assert (a.double() - r).abs().max().item() <= 2e-5 * r.abs().max().item() + 2e-6, f"{name}: {(a.double() - r).abs().max().item():.3e}"


def test_groupnorm_tokens_matches_float64(cuda):
    from latentsplat_b200.norm import group_norm_tokens
    g = torch.Generator(cuda).manual_seed(12)
    t = torch.randn(2, 100, 64, device=cuda, generator=g).requires_grad_(True)
    w = (1 + 0.2 * torch.randn(64, device=cuda, generator=g))
    b = 0.3 * torch.randn(64, device=cuda, generator=g)
    y = group_norm_tokens(t, 32, w, b, 1e-6)
    ref = F.group_norm(t.detach().double().transpose(1, 2), 32, w.double(), b.double(), 1e-6).transpose(1, 2)
    assert (y.double() - ref).abs().max().item() <= 2e-5 * ref.abs().max().item() + 2e-6
