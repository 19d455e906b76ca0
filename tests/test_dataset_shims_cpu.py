"""CPU: the input side (latentsplat_b200/dataset_shims.py) against goldens from the reference's own
crop / augmentation shims and RE10k pose conversion (tests/golden/make_golden.py::shim_goldens), plus the chunk-entry path end to
end on a synthetic JPEG chunk."""
import importlib.util
from io import BytesIO
from pathlib import Path

import numpy as np
import torch

from latentsplat_b200 import dataset_shims as ds

GOLD = Path(__file__).parent / "golden"


def _mg():
    spec = importlib.util.spec_from_file_location("make_golden", GOLD / "make_golden.py")
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def test_shims_match_reference_goldens():
    g = np.load(GOLD / "dataset_shims.npz")
    x = _mg().shim_inputs()
    img, intr = ds.rescale_and_crop(x["images"], x["intrinsics"], (64, 64))
    np.testing.assert_array_equal(img.numpy(), g["crop_image"])                  # uint8 LANCZOS path: bit-identical
    np.testing.assert_allclose(intr.numpy(), g["crop_intrinsics"], rtol=1e-7)
    refl = ds.reflect_views({"image": x["images"], "extrinsics": x["extrinsics"]})
    np.testing.assert_array_equal(refl["image"].numpy(), g["flip_image"])
    np.testing.assert_allclose(refl["extrinsics"].numpy(), g["flip_extrinsics"], rtol=1e-7, atol=1e-7)
    ex, k = ds.convert_poses(x["cameras"])
    np.testing.assert_allclose(ex.numpy(), g["pose_extrinsics"], rtol=2e-5, atol=2e-6)     # closed-form inverse vs LAPACK
    np.testing.assert_array_equal(k.numpy(), g["pose_intrinsics"])


def test_chunk_entry_to_example():
    from PIL import Image
    x = _mg().shim_inputs()
    rng = np.random.default_rng(0)
    jpegs = []
    for i in range(5):
        buf = BytesIO()
        Image.fromarray(rng.integers(0, 255, (360, 640, 3), dtype=np.uint8)).save(buf, format="JPEG", quality=95)
        jpegs.append(torch.frombuffer(bytearray(buf.getvalue()), dtype=torch.uint8))
    entry = {"key": "scene42", "cameras": x["cameras"], "images": jpegs}
    ctx, tgt = torch.tensor([0, 4]), torch.tensor([1, 2, 3])
    ex = ds.make_example(entry, ctx, tgt, (256, 256), near=1.0, far=100.0)
    assert ex["scene"] == "scene42" and ex["context"]["image"].shape == (2, 3, 256, 256) and ex["target"]["image"].shape == (3, 3, 256, 256)
    a, b = ex["context"]["extrinsics"][:, :3, 3]
    assert torch.isclose((a - b).norm(), torch.tensor(1.0), atol=1e-5)          # baseline normalised to 1
    scale = (ds.convert_poses(x["cameras"])[0][0, :3, 3] - ds.convert_poses(x["cameras"])[0][4, :3, 3]).norm()
    assert torch.allclose(ex["target"]["near"], torch.full((3,), 1.0) / scale)
    # 640x360 -> scaled to 455x256 -> centre crop 256: fx grows by 455/256 relative to the normalised value
    assert torch.allclose(ex["context"]["intrinsics"][:, 0, 0], torch.full((2,), 0.9 * 455 / 256), rtol=1e-6)
    assert ex["context"]["image"].min() >= 0 and ex["context"]["image"].max() <= 1
    # augmentation: a seeded generator flips or keeps the whole example consistently
    g = torch.Generator().manual_seed(1)
    flips = [ds.make_example(entry, ctx, tgt, (256, 256), 1.0, 100.0, augment=True, generator=g) for _ in range(6)]
    kinds = {bool(torch.equal(f["context"]["image"], ex["context"]["image"])) for f in flips}
    assert kinds == {True, False}
    flipped = next(f for f in flips if not torch.equal(f["context"]["image"], ex["context"]["image"]))
    assert torch.allclose(flipped["context"]["extrinsics"][:, 0, 3], -ex["context"]["extrinsics"][:, 0, 3], atol=1e-6)
    # a wrong image size or a degenerate baseline skips the entry
    small = BytesIO()
    Image.fromarray(np.zeros((100, 100, 3), np.uint8)).save(small, format="JPEG")
    bad = dict(entry, images=[torch.frombuffer(bytearray(small.getvalue()), dtype=torch.uint8)] * 5)
    assert ds.make_example(bad, ctx, tgt, (256, 256), 1.0, 100.0) is None
    same = dict(entry, cameras=x["cameras"][:1].repeat(5, 1))
    assert ds.make_example(same, ctx, tgt, (256, 256), 1.0, 100.0) is None
