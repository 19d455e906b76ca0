"""ctypes/numpy front end of the CPU oracle (oracle/raster_oracle.c).

TEST INFRASTRUCTURE ONLY -- see the header of raster_oracle.c.  Imported by
tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs; never by latentsplat_b200/.  PARITY UNPINNED (no reference source, tests
or golden vectors exist for the rasterizer).

The call mirrors one `GaussianRasterizer(settings)(...)` invocation of
/root/reference/src/model/decoder/cuda_splatting.py:132-158 (one view).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
_LIBS: dict[str, C.CDLL] = {}


def build(force: bool = False) -> None:
    """Compile the oracle with the committed Makefile (gcc, a few seconds)."""
    need = force or not all((_HERE / f"liboracle_{p}.so").exists() for p in ("f32", "f64"))
    if not need:
        src = (_HERE / "raster_oracle.c").stat().st_mtime
        need = any((_HERE / f"liboracle_{p}.so").stat().st_mtime < src for p in ("f32", "f64"))
    if need:
        env = dict(os.environ)
        env.pop("CC", None)
        subprocess.run(["make", "-C", str(_HERE), "-B", "all"], check=True, env=env,
                       stdout=subprocess.DEVNULL)


def _real(prec: str):
    return (C.c_float, np.float32) if prec == "f32" else (C.c_double, np.float64)


def _structs(prec: str):
    creal, _ = _real(prec)
    P = C.POINTER

    class In(C.Structure):
        _fields_ = [("G", C.c_int), ("H", C.c_int), ("W", C.c_int), ("C", C.c_int), ("sh_degree", C.c_int),
                    ("means3D", P(creal)), ("cov3D", P(creal)), ("opacity", P(creal)), ("shs", P(creal)),
                    ("colors_precomp", P(creal)), ("features", P(creal)), ("viewmatrix", P(creal)),
                    ("projmatrix", P(creal)), ("campos", P(creal)), ("tanfovx", creal), ("tanfovy", creal),
                    ("bg", P(creal)), ("scene_scale", creal)]

    class Out(C.Structure):
        _fields_ = [("depths", P(creal)), ("xy", P(creal)), ("conic_opacity", P(creal)), ("radii", P(C.c_int32)),
                    ("tiles_touched", P(C.c_uint32)), ("rgb", P(creal)), ("clamped", P(C.c_uint8)),
                    ("num_rendered", C.c_int64), ("keys_sorted", P(C.c_uint64)), ("point_list", P(C.c_uint32)),
                    ("ranges", P(C.c_uint32)), ("out_color", P(creal)), ("out_feature", P(creal)),
                    ("out_alpha", P(creal)), ("out_depth", P(creal)), ("final_T", P(creal)),
                    ("n_contrib", P(C.c_uint32)), ("flip_bound", P(creal)), ("margin_eps", creal),
                    ("marginal", P(C.c_uint8))]

    class Grad(C.Structure):
        _fields_ = [("dL_dcolor", P(creal)), ("dL_dfeature", P(creal)), ("dL_dalpha", P(creal)),
                    ("dL_ddepth", P(creal)), ("dL_dmeans3D", P(creal)), ("dL_dmeans2D", P(creal)),
                    ("dL_dshs", P(creal)), ("dL_dcolors", P(creal)), ("dL_dfeatures", P(creal)),
                    ("dL_dopacity", P(creal)), ("dL_dcov3D", P(creal)), ("dL_dconic", P(creal)),
                    ("dL_ddepths", P(creal))]

    return In, Out, Grad


def _lib(prec: str) -> C.CDLL:
    if prec not in _LIBS:
        build()
        _LIBS[prec] = C.CDLL(str(_HERE / f"liboracle_{prec}.so"))
    return _LIBS[prec]


def set_color_sh_basis(name: str) -> None:
    """'intree' (default) or '3dgs': the convention of the colour SH inside the rasterizer ([EXT], raster_oracle.c)."""
    assert name in ("intree", "3dgs")
    for prec in ("f32", "f64"):
        getattr(_lib(prec), f"oracle_set_color_sh_basis_{prec}")(int(name == "3dgs"))


def _ptr(a, ctype):
    return a.ctypes.data_as(C.POINTER(ctype)) if a is not None else None


class OracleRender:
    """Holds every array of one forward so that backward can replay it."""

    def __init__(self, **kw):
        self.__dict__.update(kw)


def forward(*, means3D, cov3D, opacity, viewmatrix, projmatrix, campos, tanfovx, tanfovy, H, W, bg=None,
            shs=None, colors_precomp=None, features=None, sh_degree=0, scene_scale=1.0, prec="f32",
            n_threads=1, margin_eps=0.0) -> OracleRender:
    creal, npreal = _real(prec)
    In, Out, _ = _structs(prec)
    lib = _lib(prec)
    f = lambda a: None if a is None else np.ascontiguousarray(np.asarray(a), dtype=npreal)
    means3D, cov3D, opacity = f(means3D), f(cov3D), f(opacity).reshape(-1)
    shs, colors_precomp, features = f(shs), f(colors_precomp), f(features)
    viewmatrix, projmatrix, campos = f(viewmatrix).reshape(-1), f(projmatrix).reshape(-1), f(campos)
    bg = f(bg if bg is not None else np.zeros(3))
    G = means3D.shape[0]
    Cf = 0 if features is None else features.shape[1]
    assert Cf <= 64
    tiles = ((H + 15) // 16) * ((W + 15) // 16)
    a = dict(depths=np.zeros(G, npreal), xy=np.zeros((G, 2), npreal), conic_opacity=np.zeros((G, 4), npreal),
             radii=np.zeros(G, np.int32), tiles_touched=np.zeros(G, np.uint32), rgb=np.zeros((G, 3), npreal),
             clamped=np.zeros((G, 3), np.uint8),
             out_color=np.zeros((3, H, W), npreal) if (shs is not None or colors_precomp is not None) else None,
             out_feature=np.zeros((Cf, H, W), npreal) if Cf else None,
             out_alpha=np.zeros((H, W), npreal), out_depth=np.zeros((H, W), npreal),
             final_T=np.zeros((H, W), npreal), n_contrib=np.zeros((H, W), np.uint32),
             flip_bound=np.zeros((H, W), npreal) if margin_eps > 0 else None,
             marginal=np.zeros(G, np.uint8) if margin_eps > 0 else None)
    i = In(G, H, W, Cf, int(sh_degree), _ptr(means3D, creal), _ptr(cov3D, creal), _ptr(opacity, creal),
           _ptr(shs, creal), _ptr(colors_precomp, creal), _ptr(features, creal), _ptr(viewmatrix, creal),
           _ptr(projmatrix, creal), _ptr(campos, creal), float(tanfovx), float(tanfovy), _ptr(bg, creal),
           float(scene_scale))
    o = Out()
    for k in ("depths", "xy", "conic_opacity", "rgb", "out_color", "out_feature", "out_alpha", "out_depth",
              "final_T", "flip_bound"):
        setattr(o, k, _ptr(a[k], creal))
    o.radii, o.tiles_touched = _ptr(a["radii"], C.c_int32), _ptr(a["tiles_touched"], C.c_uint32)
    o.clamped, o.n_contrib = _ptr(a["clamped"], C.c_uint8), _ptr(a["n_contrib"], C.c_uint32)
    o.marginal = _ptr(a["marginal"], C.c_uint8)
    o.margin_eps = float(margin_eps)
    fn = getattr(lib, f"oracle_forward_{prec}")
    fn.restype = C.c_int
    rc = fn(C.byref(i), C.byref(o), int(n_threads))
    assert rc == 0
    n = int(o.num_rendered)
    keys = np.ctypeslib.as_array(o.keys_sorted, shape=(max(n, 1),))[:n].copy()
    plist = np.ctypeslib.as_array(o.point_list, shape=(max(n, 1),))[:n].copy()
    ranges = np.ctypeslib.as_array(o.ranges, shape=(tiles, 2)).copy()
    getattr(lib, f"oracle_free_{prec}")(C.byref(o))
    return OracleRender(prec=prec, G=G, H=H, W=W, C=Cf, sh_degree=int(sh_degree), num_rendered=n, keys_sorted=keys,
                        point_list=plist, ranges=ranges, means3D=means3D, cov3D=cov3D, opacity=opacity, shs=shs,
                        colors_precomp=colors_precomp, features=features, viewmatrix=viewmatrix,
                        projmatrix=projmatrix, campos=campos, tanfovx=float(tanfovx), tanfovy=float(tanfovy),
                        bg=bg, scene_scale=float(scene_scale), **a)


def backward(r: OracleRender, *, dL_dcolor=None, dL_dfeature=None, dL_dalpha=None, dL_ddepth=None,
             n_threads=1) -> dict:
    prec = r.prec
    creal, npreal = _real(prec)
    In, Out, Grad = _structs(prec)
    lib = _lib(prec)
    f = lambda a: None if a is None else np.ascontiguousarray(np.asarray(a), dtype=npreal)
    dL_dcolor, dL_dfeature, dL_dalpha, dL_ddepth = f(dL_dcolor), f(dL_dfeature), f(dL_dalpha), f(dL_ddepth)
    G, Cf = r.G, r.C
    nsh = (r.sh_degree + 1) ** 2
    i = In(G, r.H, r.W, Cf, r.sh_degree, _ptr(r.means3D, creal), _ptr(r.cov3D, creal), _ptr(r.opacity, creal),
           _ptr(r.shs, creal), _ptr(r.colors_precomp, creal), _ptr(r.features, creal), _ptr(r.viewmatrix, creal),
           _ptr(r.projmatrix, creal), _ptr(r.campos, creal), r.tanfovx, r.tanfovy, _ptr(r.bg, creal), r.scene_scale)
    o = Out()
    for k in ("depths", "xy", "conic_opacity", "rgb", "out_color", "out_feature", "out_alpha", "out_depth",
              "final_T"):
        setattr(o, k, _ptr(getattr(r, k), creal))
    o.radii, o.tiles_touched = _ptr(r.radii, C.c_int32), _ptr(r.tiles_touched, C.c_uint32)
    o.clamped, o.n_contrib = _ptr(r.clamped, C.c_uint8), _ptr(r.n_contrib, C.c_uint32)
    o.num_rendered = r.num_rendered
    plist = np.ascontiguousarray(r.point_list if r.num_rendered else np.zeros(1, np.uint32))
    ranges = np.ascontiguousarray(r.ranges)
    o.point_list, o.ranges = _ptr(plist, C.c_uint32), _ptr(ranges, C.c_uint32)
    g = dict(dL_dmeans3D=np.zeros((G, 3), npreal), dL_dmeans2D=np.zeros((G, 2), npreal),
             dL_dshs=np.zeros((G, nsh, 3), npreal) if r.shs is not None else None,
             dL_dcolors=np.zeros((G, 3), npreal), dL_dfeatures=np.zeros((G, Cf), npreal) if Cf else None,
             dL_dopacity=np.zeros(G, npreal), dL_dcov3D=np.zeros((G, 6), npreal),
             dL_dconic=np.zeros((G, 3), npreal), dL_ddepths=np.zeros(G, npreal))
    gr = Grad(_ptr(dL_dcolor, creal), _ptr(dL_dfeature, creal), _ptr(dL_dalpha, creal), _ptr(dL_ddepth, creal),
              *[_ptr(g[k], creal) for k in ("dL_dmeans3D", "dL_dmeans2D", "dL_dshs", "dL_dcolors", "dL_dfeatures",
                                              "dL_dopacity", "dL_dcov3D", "dL_dconic", "dL_ddepths")])
    fn = getattr(lib, f"oracle_backward_{prec}")
    fn.restype = C.c_int
    rc = fn(C.byref(i), C.byref(o), C.byref(gr), int(n_threads))
    assert rc == 0
    return g


def sh_eval(deg: int, sh: np.ndarray, dirs: np.ndarray, prec="f64", order: str = "intree") -> np.ndarray:
    """sh: (N, n_ch, n_coeff) as in sh_utils.eval_sh; dirs: (N,3) unit. Returns (N, n_ch).  order='3dgs': [EXT] stock 3DGS order."""
    creal, npreal = _real(prec)
    lib = _lib(prec)
    sh = np.ascontiguousarray(np.swapaxes(np.asarray(sh, dtype=npreal), 1, 2))  # -> (N, n_coeff, n_ch)
    dirs = np.ascontiguousarray(dirs, dtype=npreal)
    N, ncoef, nch = sh.shape
    out = np.zeros((N, nch), npreal)
    fn = getattr(lib, f"oracle_sh_eval_{prec}" if order == "intree" else f"oracle_sh_eval_3dgs_{prec}")
    for n in range(N):
        fn(int(deg), int(nch), _ptr(sh[n], creal), _ptr(dirs[n], creal), _ptr(out[n], creal))
    return out
