"""GPU: the tcgen05 flash-attention core (ls_fmha_*, include/ls_fmha.h) against float64 softmax(QK^T)V.
TF32 operands (10-bit mantissa): scores are off by ~2^-10 * |q||k|, the bound below is the TF32 GEMM bound of
tests/test_gemm_gpu.py applied to both contractions."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _ref(q, k, v, scale):
    s = torch.einsum("bhld,bhmd->bhlm", q.double(), k.double()) * scale
    return torch.einsum("bhlm,bhmd->bhld", s.softmax(-1), v.double())


@pytest.mark.parametrize("B,H,L,D", [(1, 2, 128, 64), (2, 3, 1025, 64), (1, 1, 77, 64), (2, 4, 256, 128), (1, 2, 333, 128)])
def test_fmha_forward_packed_qkv(cuda, B, H, L, D):
    from latentsplat_b200.fmha import attention_packed
    g = torch.Generator(cuda).manual_seed(L + D)
    qkv = torch.randn(B, L, 3 * H * D, device=cuda, generator=g)
    out = attention_packed(qkv, H, D ** -0.5)
    q, k, v = (t.transpose(1, 2) for t in qkv.view(B, L, 3, H, D).unbind(2))           # (B, H, L, D)
    ref = _ref(q, k, v, D ** -0.5).transpose(1, 2).reshape(B, L, H * D)
    err = (out.double() - ref).abs().max().item()
    assert err <= 2e-2 * ref.abs().max().item() + 1e-3, f"max err {err:.3e} (ref max {ref.abs().max().item():.3e})"


@pytest.mark.parametrize("B,H,L,D", [(1, 2, 128, 64), (2, 3, 1025, 64), (1, 1, 77, 64), (2, 4, 256, 128), (1, 2, 333, 128)])
def test_fmha_backward_packed_qkv(cuda, B, H, L, D):
    """dq, dk, dv of the two backward kernels (dq per query block; dk/dv per key block, transposed) against float64 autograd."""
    from latentsplat_b200.fmha import attention_packed
    g = torch.Generator(cuda).manual_seed(7 * L + D)
    qkv = torch.randn(B, L, 3 * H * D, device=cuda, generator=g).requires_grad_(True)
    go = torch.randn(B, L, H * D, device=cuda, generator=g)
    out = attention_packed(qkv, H, D ** -0.5)
    (grad,) = torch.autograd.grad(out, qkv, go)

    ref_in = qkv.detach().double().requires_grad_(True)
    q, k, v = (t.transpose(1, 2) for t in ref_in.view(B, L, 3, H, D).unbind(2))
    ref = _ref(q, k, v, D ** -0.5).transpose(1, 2).reshape(B, L, H * D)
    (ref_grad,) = torch.autograd.grad(ref, ref_in, go.double())
    for name, ours, want in zip("qkv", grad.view(B, L, 3, H * D).unbind(2), ref_grad.view(B, L, 3, H * D).unbind(2)):
        err = (ours.double() - want).abs().max().item()
        bound = 2e-2 * want.abs().max().item() + 1e-3
        assert err <= bound, f"d{name}: max err {err:.3e} > {bound:.3e}"
    assert torch.isfinite(grad).all()


@pytest.mark.parametrize("b,L,c", [(2, 1024, 512), (1, 256, 128), (1, 64, 32)])
def test_wide_single_head_attention(cuda, b, L, c):
    """VAE mid-block attention: scores through ls_gemm_tf32 + in-place row softmax kernels, forward and all gradients."""
    from latentsplat_b200.fmha import attention_wide
    g = torch.Generator(cuda).manual_seed(L + c)
    q, k, v = (torch.randn(b, L, c, device=cuda, generator=g).requires_grad_(True) for _ in range(3))
    go = torch.randn(b, L, c, device=cuda, generator=g)
    out = attention_wide(q, k, v, c ** -0.5)
    grads = torch.autograd.grad(out, (q, k, v), go)
    qd, kd, vd = (t.detach().double().requires_grad_(True) for t in (q, k, v))
    ref = _ref(qd[:, None], kd[:, None], vd[:, None], c ** -0.5)[:, 0]
    ref_grads = torch.autograd.grad(ref, (qd, kd, vd), go.double())
    for name, ours, want in zip(("out", "dq", "dk", "dv"), (out, *grads), (ref, *ref_grads)):
        err = (ours.double() - want).abs().max().item()
        bound = 2e-2 * want.abs().max().item() + 1e-3
        assert err <= bound, f"{name}: max err {err:.3e} > {bound:.3e}"
