"""CPU: the C-ABI library builds for sm_100a, loads and exports every symbol include/ls_raster.h declares;
argument validation works without a GPU (no kernel is launched here)."""
import ctypes as C
import re
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]


def test_library_exports_every_declared_symbol(lib):
    header = (ROOT / "include" / "ls_raster.h").read_text()
    both = "".join(p.read_text() for p in sorted((ROOT / "include").glob("*.h")))      # every header of the C ABI
    declared = set(re.findall(r"LS_API\s+[\w\s\*]+?\b(ls_\w+)\s*\(", both))
    from latentsplat_b200 import _capi
    assert declared == set(_capi.EXPORTS), declared
    for name in declared:
        assert hasattr(lib, name), f"{name} not exported"
    assert lib.ls_raster_abi_version() == _capi.ABI_VERSION == int(re.search(r"LS_RASTER_ABI_VERSION (\d+)", header).group(1))


def test_struct_layouts_match_the_header():
    """Field order of the ctypes mirrors == field order in the header (names), a cheap drift detector."""
    from latentsplat_b200 import _capi
    header = (ROOT / "include" / "ls_raster.h").read_text()
    header = "".join(p.read_text() for p in sorted((ROOT / "include").glob("*.h")))
    for name in ("LsRasterScene", "LsRasterState", "LsRasterImages", "LsRasterGrads", "LsRasterSizes", "LsGemmArgs",
                 "LsEpipolarGather", "LsGroupNorm", "LsConv2d"):
        body = re.search(r"typedef struct %s \{(.*?)\} %s;" % (name, name), header, re.S).group(1)
        body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
        fields = []
        for decl in body.split(";"):
            decl = decl.strip()
            if not decl:
                continue
            names = decl.split(None, 1)[1] if not decl.startswith("const") else decl.split(None, 2)[2]
            for n in names.split(","):
                fields.append(n.replace("*", "").strip())
        assert fields == [f[0] for f in getattr(_capi, name)._fields_], name


def test_sizes_and_validation_without_gpu(lib):
    from latentsplat_b200 import _capi
    sc = _capi.LsRasterScene()
    sc.n_views, sc.views_per_scene, sc.G, sc.H, sc.W, sc.C = 4, 2, 1000, 250, 130, 4
    sc.color_mode, sc.feature_mode = _capi.COLOR_SH, _capi.FEATURE_PRECOMP
    sz = _capi.LsRasterSizes()
    assert lib.ls_raster_sizes(C.byref(sc), C.byref(sz)) == 0
    assert (sz.n_scenes, sz.tiles_per_view, sz.chan_stride, sz.grad_stride) == (2, 16 * 9, 8, 16)
    assert sz.tile_slots == 4 * 16 * 9 and sz.pixels == 4 * 250 * 130 and sz.n_value_channels == 7
    # invalid scenes are rejected before anything touches the device
    st, im = _capi.LsRasterState(), _capi.LsRasterImages()
    sc.views_per_scene = 3
    assert lib.ls_raster_forward(C.byref(sc), C.byref(st), C.byref(im), 3, None) != 0
    assert b"views_per_scene" in lib.ls_last_error()
    sc.views_per_scene = 2
    sc.color_mode, sc.feature_mode, sc.C = _capi.COLOR_NONE, _capi.FEATURE_NONE, 0
    assert lib.ls_raster_forward(C.byref(sc), C.byref(st), C.byref(im), 3, None) != 0
    assert b"nothing to render" in lib.ls_last_error()
    sc.feature_mode, sc.C = _capi.FEATURE_PRECOMP, 40
    assert lib.ls_raster_forward(C.byref(sc), C.byref(st), C.byref(im), 3, None) != 0
    assert b"LS_MAX_VALUE_CHANNELS" in lib.ls_last_error()


def test_conv_descriptor_validation_without_gpu(lib):
    """ls_conv2d_*: sizes and argument checks run on the host before anything is enqueued."""
    from latentsplat_b200 import _capi
    d = _capi.LsConv2d(2, 64, 48, 32, 64, 3, 3, 1, 1, 0)
    oh, ow = C.c_int32(), C.c_int32()
    assert lib.ls_conv2d_out_size(C.byref(d), C.byref(oh), C.byref(ow)) == 0 and (oh.value, ow.value) == (64, 48)
    d = _capi.LsConv2d(2, 64, 64, 32, 64, 4, 4, 2, 1, 0)
    assert lib.ls_conv2d_out_size(C.byref(d), C.byref(oh), C.byref(ow)) == 0 and (oh.value, ow.value) == (32, 32)
    d = _capi.LsConv2d(2, 16, 16, 32, 64, 4, 4, 4, 0, 1)
    assert lib.ls_conv2d_out_size(C.byref(d), C.byref(oh), C.byref(ow)) == 0 and (oh.value, ow.value) == (64, 64)
    bad = _capi.LsConv2d(2, 16, 16, 3, 64, 3, 3, 1, 1, 0)                      # channels must be padded to a multiple of 4
    assert lib.ls_conv2d_out_size(C.byref(bad), C.byref(oh), C.byref(ow)) != 0 and b"multiples of 4" in lib.ls_last_error()
    bad = _capi.LsConv2d(2, 16, 16, 32, 64, 3, 3, 2, 0, 1)                     # overlapping transposed convolution
    assert lib.ls_conv2d_forward(C.byref(bad), None, None, None, None, None, 0, None) != 0 and b"transposed" in lib.ls_last_error()
    assert lib.ls_conv2d_forward(C.byref(d), None, None, None, None, None, 0, None) != 0 and b"NULL" in lib.ls_last_error()


def test_product_fails_loudly_without_cuda():
    """No CPU fallback: CPU tensors are rejected, never silently rendered by something else."""
    import torch
    from diff_gaussian_rasterization import GaussianRasterizationSettings, GaussianRasterizer
    s = GaussianRasterizationSettings(16, 16, 0.5, 0.5, torch.zeros(3), 1.0, torch.eye(4), torch.eye(4), 0,
                                      torch.zeros(3), False, False)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        GaussianRasterizer(s)(means3D=torch.zeros(4, 3), means2D=None, opacities=torch.ones(4, 1),
                              colors_precomp=torch.zeros(4, 3), cov3D_precomp=torch.zeros(4, 6))


def test_product_never_imports_the_oracle():
    pkg = ROOT / "latentsplat_b200"
    for p in list(pkg.rglob("*.py")) + list(pkg.rglob("*.cu")) + list(pkg.rglob("*.cuh")) + list((ROOT / "diff_gaussian_rasterization").rglob("*.py")):
        txt = p.read_text()
        assert not re.search(r"^\s*(from|import)\s+oracle\b", txt, re.M), p
        assert "liboracle" not in txt, p
