"""CPU: inference callers and formats (latentsplat_b200/inference.py): Lightning checkpoint loading by the
reference's parameter names, the evaluation-index JSON, test_step's view chunking and PNG layout."""
import json
from pathlib import Path

import numpy as np
import pytest
import torch

from latentsplat_b200 import inference


def _tiny_pipe():
    from latentsplat_b200.configs import build_modules
    from latentsplat_b200.pipeline import RenderPipeline
    torch.manual_seed(0)
    ae, enc, dec, disc = build_modules(with_discriminator=True)
    return RenderPipeline(ae, enc, dec, disc)


def test_lightning_checkpoint_round_trip_by_reference_names(tmp_path):
    pipe = _tiny_pipe()
    # a ModelWrapper-style state dict: sub-module prefixes + wrapper-only entries that must be ignored
    state = {f"{n}.{k}": v.clone() for n in ("encoder", "autoencoder", "discriminator") for k, v in getattr(pipe, n).state_dict().items()}
    state["target_combined_losses.nll_losses.1.lpips.lpips.lin0.model.1.weight"] = torch.zeros(1, 64, 1, 1)
    assert "encoder.epipolar_transformer.transformer.layers.0.0.fn.to_kv.weight" in state
    assert "autoencoder.model.decoder.up_blocks.0.resnets.0.conv1.weight" in state and "encoder.backbone.dino.blocks.0.attn.qkv.weight" in state
    for v in state.values():
        if v.dtype.is_floating_point:
            v.add_(1.0)
    path = tmp_path / "last.ckpt"
    torch.save({"state_dict": state, "global_step": 7}, path)
    fresh = _tiny_pipe()
    report = inference.load_lightning_checkpoint(path, fresh)
    assert report["missing"] == [] and report["unexpected"] == [] and len(report["ignored"]) == 1
    for n in ("encoder", "autoencoder", "discriminator"):
        for k, v in getattr(fresh, n).state_dict().items():
            assert torch.equal(v, state[f"{n}.{k}"]), f"{n}.{k}"
    bad = dict(state)
    bad.pop("encoder.to_gaussians.1.weight")
    with pytest.raises(RuntimeError, match="missing"):
        inference.load_lightning_checkpoint({"state_dict": bad}, _tiny_pipe())
    assert inference.load_lightning_checkpoint({"state_dict": bad}, _tiny_pipe(), strict=False)["missing"] == ["encoder.to_gaussians.1.weight"]


def test_evaluation_index_format(tmp_path):
    raw = {"sceneA": [{"context": [16, 40], "target": list(range(20))}, {"context": [16, 40], "target": [20, 21]}],
           "sceneB": None, "sceneC": {"context": [3, 9], "target": [4, 5, 6]}}
    p = tmp_path / "index.json"
    p.write_text(json.dumps(raw))
    idx = inference.EvaluationIndex(p)
    assert len(idx) == 3 and [s for s, _ in idx] == ["sceneA", "sceneA", "sceneC"]
    assert idx.entries("sceneA")[1].target == (20, 21) and idx.entries("sceneB") == [] and idx.entries("sceneC")[0].context == (3, 9)


def test_png_layout_and_quantisation(tmp_path):
    from PIL import Image
    imgs = torch.stack([torch.full((3, 4, 6), 0.5), torch.linspace(-0.2, 1.2, 72).view(3, 4, 6)])
    paths = inference.save_predictions(imgs, [7, 123], tmp_path, "scene0", [40, 16])
    assert [p.relative_to(tmp_path).as_posix() for p in paths] == ["scene0/16_40/color/000007.png", "scene0/16_40/color/000123.png"]
    back = np.asarray(Image.open(paths[0]))
    assert back.shape == (4, 6, 3) and (back == 127).all()                      # 0.5 * 255 truncated, as the reference's .type(uint8)
    back = np.asarray(Image.open(paths[1]))
    assert back.min() == 0 and back.max() == 255


def test_chunked_prediction_runs_the_encoder_once_and_concatenates_views():
    """predict_target_views on stand-in modules: chunks of views through decoder / VAE, deterministic mode uses the modes."""
    from types import SimpleNamespace
    calls = {"encoder": 0, "decoder": []}

    class G:
        def sample(self): return "sampled"
        def mode(self): return "mode"
        def flatten(self): return "flat"

    class Post:
        def __init__(self, x): self.x = x
        def sample(self): return self.x + 1.0
        def mode(self): return self.x

    def encoder(ctx, step, features=None, deterministic=False):
        calls["encoder"] += 1
        return G()

    def decoder(g, ex, intr, near, far, size):
        calls["decoder"].append((g, ex.shape[1]))
        v = ex.shape[1]
        base = ex[:, :, 0, 0].view(1, v, 1, 1, 1).expand(1, v, 4, 16, 16).clone()
        return SimpleNamespace(feature_posterior=Post(base), color=torch.zeros(1, v, 3, 16, 16))

    ae = SimpleNamespace(expects_skip=False, expects_skip_extra=False, decode=lambda z, skip: z[:, :, :3].repeat_interleave(8, -1).repeat_interleave(8, -2))
    pipe = SimpleNamespace(encoder=encoder, decoder=decoder, autoencoder=ae, variational="gaussians", supersampling_factor=8)
    v = 5
    ex = torch.eye(4).repeat(1, v, 1, 1)
    ex[0, :, 0, 0] = torch.arange(v, dtype=torch.float32)
    batch = {"context": {"image": torch.zeros(1, 2, 3, 16, 16)},
             "target": {"extrinsics": ex, "intrinsics": torch.eye(3).repeat(1, v, 1, 1), "near": torch.ones(1, v), "far": torch.ones(1, v)}}
    out = inference.predict_target_views(pipe, batch, views_per_chunk=2)
    assert calls["encoder"] == 1 and calls["decoder"] == [("sampled", 2), ("sampled", 2), ("sampled", 1)]
    assert out.shape == (v, 3, 16, 16) and torch.allclose(out[:, 0, 0, 0], torch.arange(v, dtype=torch.float32) + 1.0)
    det = inference.predict_target_views(pipe, batch, deterministic=True)
    assert calls["decoder"][-1] == ("mode", v) and torch.allclose(det[:, 0, 0, 0], torch.arange(v, dtype=torch.float32))
    with pytest.raises(ValueError):
        inference.predict_target_views(pipe, {"context": {"image": torch.zeros(2, 2, 3, 16, 16)}, "target": batch["target"]})
