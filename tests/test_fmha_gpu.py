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
