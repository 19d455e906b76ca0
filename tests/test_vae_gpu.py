"""GPU: the kl-f8 VAE decoder (`AutoencoderKL.decode`, autoencoder_kl.py:93-124) with every sm_100a kernel on -- tcgen05
implicit-GEMM 3x3 / 1x1 convolutions in NHWC, the up-sampling folded into the 3x3 filter, NHWC GroupNorm+SiLU, the mid-block
attention through ls_gemm_tf32 + row softmax -- against a float64 CPU run of the SAME module (same parameters).

What this pins: the kernels against the module's definition.  What it cannot pin: the module's definition against diffusers
0.25.1 itself (not installed here; the restated details are listed in DESIGN.md "diffusers details restated").
Tolerance (stated, TF32): operands rounded to 10-bit mantissas, fp32 accumulation, 30 convolutions deep with GroupNorm
renormalising in between: outputs within 2e-2 of the output range, gradients within 5e-2 of each tensor's largest entry.
"""
import copy

import pytest
import torch

pytestmark = pytest.mark.gpu


def test_vae_decode_matches_float64_module(cuda):
    from latentsplat_b200 import _capi
    from latentsplat_b200.model.autoencoder.vae_kl import AutoencoderKLModel
    torch.manual_seed(11)
    ref = AutoencoderKLModel().double()                        # kl-f8: (128, 256, 512, 512), 2 layers per block, 4 latent channels
    for name, p in ref.named_parameters():                     # non-trivial norm affine parameters
        if "norm" in name:
            p.data.add_(0.1 * torch.randn_like(p))
    ours = copy.deepcopy(ref).float().to(cuda).to(memory_format=torch.channels_last)
    z64 = torch.randn(1, 4, 16, 16, dtype=torch.float64, requires_grad=True)
    z = z64.detach().float().to(cuda).requires_grad_(True)
    w64 = torch.randn(1, 3, 128, 128, dtype=torch.float64)

    launches = _capi.KERNEL_LAUNCHES[0]
    out = ours.decode(z)
    assert _capi.KERNEL_LAUNCHES[0] - launches > 60, "decode did not run on our kernels"
    probe = [ours.decoder.conv_in.weight, ours.decoder.mid_block.attentions[0].to_q.weight,
             ours.decoder.up_blocks[1].resnets[0].conv1.weight, ours.decoder.up_blocks[0].upsamplers[0].conv.weight,
             ours.decoder.conv_out.weight, ours.decoder.conv_norm_out.weight]
    grads = torch.autograd.grad((out * w64.float().to(cuda)).sum(), [z] + probe)

    out64 = ref.decode(z64)
    probe64 = [ref.decoder.conv_in.weight, ref.decoder.mid_block.attentions[0].to_q.weight,
               ref.decoder.up_blocks[1].resnets[0].conv1.weight, ref.decoder.up_blocks[0].upsamplers[0].conv.weight,
               ref.decoder.conv_out.weight, ref.decoder.conv_norm_out.weight]
    grads64 = torch.autograd.grad((out64 * w64).sum(), [z64] + probe64)

    assert out.shape == (1, 3, 128, 128)
    err = (out.double().cpu() - out64).abs().max().item()
    rng = out64.abs().max().item()
    assert err <= 2e-2 * rng, f"decode: max err {err:.3e} vs range {rng:.3e}"
    names = ["z", "conv_in.w", "mid.attn.to_q.w", "up1.res0.conv1.w", "up0.upsample.conv.w", "conv_out.w", "norm_out.w"]
    for name, g, g64 in zip(names, grads, grads64):
        e = (g.double().cpu() - g64).abs().max().item()
        s = g64.abs().max().item()
        assert e <= 5e-2 * s, f"d/d{name}: max err {e:.3e} vs largest entry {s:.3e}"
