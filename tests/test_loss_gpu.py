"""GPU: the loss heads on CUDA -- LPIPS-VGG on the tcgen05 implicit-GEMM convolutions and the adaptive GAN weight through the VAE
decoder's last layer (two retained partial backward passes through our convolution autograd functions) -- against the same
modules on the CPU in fp32.  Tolerance: TF32 operands through 13 convolutions -> 2e-2 relative on the scalar values."""
import copy

import pytest
import torch

pytestmark = pytest.mark.gpu


def test_lpips_vgg_on_cuda_matches_cpu(cuda):
    from latentsplat_b200 import _capi
    from latentsplat_b200.loss import LpipsVgg
    torch.manual_seed(2)
    ref = LpipsVgg(weights="random")
    ours = copy.deepcopy(ref).to(cuda).to(memory_format=torch.channels_last)
    a, b = torch.rand(2, 3, 64, 64), torch.rand(2, 3, 64, 64)
    a_gpu = a.to(cuda).requires_grad_(True)
    n0 = _capi.KERNEL_LAUNCHES[0]
    got = ours(a_gpu, b.to(cuda))
    (g_gpu,) = torch.autograd.grad(got, a_gpu)
    assert _capi.KERNEL_LAUNCHES[0] - n0 >= 26, "LPIPS did not run on our convolution kernels"
    a_cpu = a.clone().requires_grad_(True)
    want = ref(a_cpu, b)
    (g_cpu,) = torch.autograd.grad(want, a_cpu)
    assert abs(got.item() - want.item()) <= 2e-2 * abs(want.item()), (got.item(), want.item())
    # the image gradient of a randomly initialised 13-layer VGG is ~1e-6 per pixel and passes 26 TF32 contractions (forward and
    # input-gradient): compare direction and size of the whole field, not single pixels
    g_gpu = g_gpu.cpu().double().flatten()
    g_ref = g_cpu.double().flatten()
    cos = torch.dot(g_gpu, g_ref) / (g_gpu.norm() * g_ref.norm())
    rel = (g_gpu - g_ref).norm() / g_ref.norm()
    print(f"d lpips / d image: cosine {cos.item():.5f}, relative L2 error {rel.item():.3e}")
    assert torch.isfinite(g_gpu).all() and cos.item() > 0.999 and rel.item() < 5e-2      # measured on B200: 0.99982, 1.9e-2


def test_adaptive_gan_weight_through_decoder_last_layer(cuda):
    """LossGroup.forward_generator with last_layer_weights = conv_out.weight of a small VAE decoder: the two torch.autograd.grad
    calls (retain_graph) run through our convolution / GroupNorm autograd functions; value == the CPU fp32 run."""
    from latentsplat_b200.loss import LossGeneratorCfg, LossGroupCfg, LossL1Cfg, LossMseCfg, get_loss_group
    from latentsplat_b200.model.autoencoder.vae_kl import AutoencoderKLModel
    from latentsplat_b200.model.types import GroundTruth, Prediction
    torch.manual_seed(5)
    vae = AutoencoderKLModel(block_out_channels=(32, 64), layers_per_block=1, latent_channels=4, norm_num_groups=8)
    disc = torch.nn.Conv2d(3, 1, 4, stride=4)
    z = torch.randn(2, 4, 16, 16)
    target = torch.rand(1, 2, 3, 32, 32)
    group = get_loss_group("combined", LossGroupCfg(nll=[LossMseCfg(weight=10), LossL1Cfg()], generator=LossGeneratorCfg(weight=0.5)))
    res = []
    for dev in ("cpu", cuda):
        m = copy.deepcopy(vae).to(dev)
        d = copy.deepcopy(disc).to(dev)
        if dev != "cpu":
            m = m.to(memory_format=torch.channels_last)
        image = m.decode(z.to(dev))[None]
        pred = Prediction(image=image, logits_fake=d(image[0])[None])
        total, parts = group.forward_generator(pred, GroundTruth(image=target.to(dev)), 0, last_layer_weights=m.decoder.conv_out.weight)
        total.backward()
        gen = parts["combined/generator"]
        res.append((total.item(), gen.weighted.item() / (0.5 * gen.unweighted.item()),       # = the adaptive weight
                    m.decoder.conv_in.weight.grad.detach().cpu()))
    (t0, w0, g0), (t1, w1, g1) = res
    assert 0.0 < w0 <= 1.0
    assert abs(t1 - t0) <= 2e-2 * abs(t0), (t0, t1)
    assert abs(w1 - w0) <= 3e-2 * max(abs(w0), 1e-3), (w0, w1)
    assert (g1 - g0).abs().max().item() <= 5e-2 * g0.abs().max().item()
