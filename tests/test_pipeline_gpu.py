"""GPU: the whole hot path (encoder -> CUDA splat -> VAE decode) against the same modules on the CPU with the
oracle rasterizer behind the reference's per-view data flow (oracle/decoder_cpu.py).
Tolerance: the CUDA path runs every Linear on tcgen05 kind::tf32 (operands truncated to 10 mantissa bits) and
its convolutions with cuDNN's TF32 default, the CPU path is fp32, and the encoder's positional encodings reach
frequency 2 pi 512: outputs agree to ~1e-2 of their range, which is what is asserted."""
import importlib.util
from fractions import Fraction
from pathlib import Path

import numpy as np
import pytest
import torch

import helpers

pytestmark = pytest.mark.gpu
GOLD = Path(__file__).parent / "golden"


def _build(decoder):
    from latentsplat_b200.model.autoencoder import AutoencoderKLCfg, get_autoencoder
    from latentsplat_b200.pipeline import RenderPipeline
    from test_encoder_cpu import build_encoder
    cfg = AutoencoderKLCfg("kl", "kl_f8", ["DownEncoderBlock2D"] * 4, ["UpDecoderBlock2D"] * 4, [32, 64, 64, 64], 1, 4,
                           True, True, True, False)
    ae = get_autoencoder(cfg, 3, 3, 32)
    enc = build_encoder(variational=True, n_feature_channels=4)
    pipe = RenderPipeline(ae, enc, decoder, None, supersampling_factor=8)
    helpers.init_by_name(pipe, seed=21)
    return pipe.eval()


def _batch(hw=64):
    spec = importlib.util.spec_from_file_location("make_golden", GOLD / "make_golden.py")
    mg = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mg)
    ctx = mg.encoder_context(b=1, v=2, hw=hw, seed=9)
    from latentsplat_b200 import synthetic
    tgt = {"extrinsics": synthetic.pose(0.3, -3.0, 0.02, 0.0)[None, None], "intrinsics": ctx["intrinsics"][:, :1],
           "near": ctx["near"][:, :1], "far": ctx["far"][:, :1],
           "image": torch.rand(1, 1, 3, hw, hw, generator=torch.Generator().manual_seed(2))}
    return {"context": ctx, "target": tgt}


def test_full_pipeline_matches_cpu_reference_flow(cuda):
    from latentsplat_b200.model.decoder import DecoderSplattingCUDA, DecoderSplattingCUDACfg
    from oracle.decoder_cpu import DecoderSplattingCPU
    batch = _batch()
    cpu = _build(DecoderSplattingCPU((0.1, 0.2, 0.3)))
    with torch.no_grad():
        ref = cpu(batch, deterministic=True)
    gpu = _build(DecoderSplattingCUDA(DecoderSplattingCUDACfg("splatting_cuda"), [0.1, 0.2, 0.3])).to(cuda)
    dev_batch = {k: {n: t.to(cuda) for n, t in v.items()} for k, v in batch.items()}
    out = gpu(dev_batch, deterministic=True)

    def close(a, b, name, tol):
        a, b = a.detach().cpu().numpy(), b.numpy()
        err = np.abs(a - b)
        scale = np.abs(b).max()
        # deterministic mode picks the arg-max depth bucket per ray: TF32 noise flips near-ties, moving a handful of
        # Gaussians to a different bucket (a discontinuity of the model, not of the kernels) -> quantile + mean bounds
        assert np.quantile(err, 0.98) <= tol * scale, f"{name}: p98 err {np.quantile(err, 0.98):.3e} vs scale {scale:.3e}"
        assert err.mean() <= tol * scale, f"{name}: mean err {err.mean():.3e} vs scale {scale:.3e}"
    close(out.gaussians.means, ref.gaussians.means, "gaussian means", 1e-2)
    close(out.render.color, ref.render.color, "rendered colour", 2e-2)
    close(out.render.mask, ref.render.mask, "mask", 2e-2)
    close(out.z, ref.z, "latent z", 3e-2)
    close(out.image, ref.image, "decoded image", 3e-2)
    assert out.image.shape == (1, 1, 3, 64, 64) and out.z.shape == (1, 1, 4, 8, 8)

    # backward through everything (stochastic path), gradients finite and non-trivial
    out = gpu(dev_batch, deterministic=False)
    loss = 10 * ((out.render.color - dev_batch["target"]["image"]) ** 2).mean() + (out.image - dev_batch["target"]["image"]).abs().mean()
    loss.backward()
    for name in ("encoder.backbone.dino.blocks.0.attn.qkv.weight", "encoder.epipolar_transformer.transformer.layers.0.0.fn.to_kv.weight",
                 "encoder.to_gaussians.1.weight", "autoencoder.model.decoder.conv_in.weight", "autoencoder.skip_convs.0.weight"):
        g = dict(gpu.named_parameters())[name].grad
        assert g is not None and torch.isfinite(g).all() and float(g.abs().max()) > 0, name
