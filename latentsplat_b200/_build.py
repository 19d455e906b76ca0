"""In-tree build of libls_raster.so with nvcc for sm_100a (no torch headers involved).

`python -m latentsplat_b200._build` or `__graft_entry__.build()`.  The .so lands in
latentsplat_b200/lib/ (git-ignored, but it travels to the GPU box with the snapshot).
"""
from __future__ import annotations

import os
import shutil
import subprocess
from pathlib import Path

PKG = Path(__file__).resolve().parent
ROOT = PKG.parent
CSRC = PKG / "csrc"
LIB_DIR = PKG / "lib"
LIB = LIB_DIR / "libls_raster.so"

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", shutil.which("nvcc")):
        if cand and Path(cand).exists():
            return cand
    raise RuntimeError("nvcc not found; libls_raster.so cannot be built")


def sources() -> list[Path]:
    return sorted(CSRC.glob("*.cu"))


def is_stale() -> bool:
    if not LIB.exists():
        return True
    t = LIB.stat().st_mtime
    deps = list(CSRC.glob("*")) + list((ROOT / "include").glob("*.h"))
    return any(d.stat().st_mtime > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> Path:
    """Build if stale.  Safe to call from several processes at once (one rank per GPU under torchrun): an exclusive file
    lock serialises them, the link goes to a temporary name and is renamed into place, late-comers find a fresh library."""
    if not force and not is_stale():
        return LIB
    import fcntl
    LIB_DIR.mkdir(exist_ok=True)
    with open(LIB_DIR / ".build.lock", "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            if not force and not is_stale():
                return LIB
            return _build_locked(verbose)
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)


def _build_locked(verbose: bool) -> Path:
    obj_dir = LIB_DIR / "obj"
    obj_dir.mkdir(exist_ok=True)
    nvcc = _nvcc()
    env = dict(os.environ)
    env.pop("CC", None)
    env.pop("CXX", None)
    procs = []
    objs = []
    for src in sources():
        obj = obj_dir / (src.stem + ".o")
        objs.append(obj)
        cmd = [nvcc, *NVCC_FLAGS, "-ccbin", "/usr/bin/g++" if Path("/usr/bin/g++").exists() else "g++",
               f"-I{ROOT / 'include'}", f"-I{CSRC}", "-c", str(src), "-o", str(obj)]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        procs.append((src, subprocess.Popen(cmd, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    for src, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src.name}:\n{out}")
        if verbose and out:
            print(out)
    tmp = LIB.with_suffix(f".so.tmp{os.getpid()}")
    link = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-ccbin",
            "/usr/bin/g++" if Path("/usr/bin/g++").exists() else "g++", "-o", str(tmp),
            *map(str, objs), "-lcudart"]
    r = subprocess.run(link, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}")
    os.replace(tmp, LIB)
    return LIB


if __name__ == "__main__":
    import sys
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
