"""GPU: the whole encoder with every sm_100a kernel on (tcgen05 GEMMs / convolutions / flash attention, fused epipolar sampler,
fused depth-sampling + Gaussian-adapter tail) against tests/golden/encoder.npz, which the REFERENCE's own PyTorch modules
produced in fp32 on the CPU (tests/golden/make_golden.py::encoder_goldens).

Tolerance (stated, TF32): every contraction on the GPU path rounds its operands to TF32 (10-bit mantissa, 2^-11 relative per
operand) and accumulates in fp32; through the 12 DINO blocks + 2 epipolar layers + heads that compounds to ~1e-3..1e-2 of an
activation's range, so a Gaussian counts as matching when all of its values are within 2e-2 of the tensor's range.  The
deterministic mode picks the arg-max depth bucket per ray (depth_predictor_monocular.py:58-60): where the two best buckets of the
fp32 reference are closer than the TF32 noise the GPU may pick the other one and that Gaussian moves as a whole (another depth).
Those are counted separately and bounded (<= 5 % of rays); means / covariances of non-flipped rays must match.
"""
import numpy as np
import pytest
import torch

import helpers
from test_encoder_cpu import GOLD, _mg, build_encoder

pytestmark = pytest.mark.gpu


def test_gpu_encoder_matches_reference_golden(cuda, monkeypatch):
    g = np.load(GOLD / "encoder.npz")
    model = build_encoder()
    helpers.init_by_name(model, seed=3)
    model = model.to(cuda).to(memory_format=torch.channels_last)
    ctx = {k: v.to(cuda) for k, v in _mg().encoder_context().items()}
    from latentsplat_b200 import _capi, gaussian_head
    tail_calls = []
    fused = gaussian_head.gaussian_head
    monkeypatch.setattr(gaussian_head, "gaussian_head", lambda *a, **k: (tail_calls.append(1), fused(*a, **k))[1])
    launches = _capi.KERNEL_LAUNCHES[0]
    with torch.no_grad():
        det = model(ctx, 0, deterministic=True)
    assert _capi.KERNEL_LAUNCHES[0] - launches > 100, "the encoder did not run on our kernels"
    assert tail_calls, "the fused depth-sampling + Gaussian-adapter tail (k_ghead_*) was bypassed"     # a dead guard hid this once
    got = dict(means=det.means, cov=det.covariances, opac=det.opacities, csh=det.color_harmonics, fsh=det.feature_harmonics.params)
    want_means = g["det_means"]
    a_means = got["means"].cpu().numpy()[:, ::5]
    # a flipped depth bucket moves the mean along its ray by at least one bucket (1/32 of the disparity range)
    moved = np.abs(a_means - want_means).max(axis=-1) > 2e-2 * np.abs(want_means).max()
    assert moved.mean() <= 0.05, f"{moved.sum()} of {moved.size} rays picked another depth bucket"
    worst = {}
    for k, v in got.items():
        want = g[f"det_{k}"]
        a = v.cpu().numpy()[:, ::5]
        assert a.shape == want.shape and np.isfinite(a).all()
        scale = np.abs(want).max()
        err = np.abs(a - want).reshape(a.shape[0], a.shape[1], -1).max(axis=-1)
        err = np.where(moved, 0.0, err)
        worst[k] = float(err.max() / scale)
        assert err.max() <= 2e-2 * scale, f"det_{k}: max err {err.max():.3e} vs range {scale:.3e} (TF32 bound 2e-2)"
    print("encoder GPU vs golden, max err / range:", {k: f"{v:.2e}" for k, v in worst.items()}, "flipped rays:", int(moved.sum()))
