#!/usr/bin/env python
"""Per-kernel SASS mnemonic summary of latentsplat_b200/lib/libls_raster.so (cuobjdump -sass), written to
profiles/r02_sass_{gemm,conv,fmha,raster}.txt: for every kernel the instruction count, the counts of the mnemonics that show
which hardware path it uses (UTCHMMA = tcgen05.mma, UTMALDG = TMA tensor load, UBLKCP = cp.async.bulk, LDTM = tcgen05.ld,
UTCBAR = tcgen05.commit, SYNCS = mbarrier, RED = red.global, LDGSTS = cp.async, LDL/STL = local memory), and the first lines
that carry them.

    python scripts/sass_summary.py            # needs cuobjdump + c++filt (CUDA toolkit / binutils), no GPU
"""
import collections
import pathlib
import re
import subprocess

ROOT = pathlib.Path(__file__).resolve().parents[1]
LIB = ROOT / "latentsplat_b200" / "lib" / "libls_raster.so"
SHOW = ("UTCHMMA", "UTCQMMA", "UTMALDG", "UBLKCP", "LDTM", "UTCBAR", "RED", "REDG", "LDGSTS")       # counted with modifiers + listed
COUNT = SHOW + ("SYNCS", "ATOMG", "HMMA", "FFMA", "MUFU", "LDL", "STL", "SHFL", "VOTE", "BAR", "LDG", "STG", "LDS", "STS")
FAMILIES = {"gemm": ("ls_gemm", "ls_attn", "ls_norm"), "conv": ("ls_conv", "ls_norm_nhwc"), "fmha": ("ls_fmha",),
            "raster": ("ls_raster_fwd", "ls_raster_bwd", "ls_epipolar", "ls_ghead")}


def main():
    dump = subprocess.run(["cuobjdump", "-sass", str(LIB)], capture_output=True, text=True, check=True).stdout
    kernels, src, cur = collections.OrderedDict(), None, None
    for line in dump.splitlines():
        m = re.match(r"identifier = (\S+)", line.strip())
        if m:
            src = pathlib.Path(m.group(1)).stem
            continue
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = (src, m.group(1))
            kernels[cur] = []
        elif cur and re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+\S", line):
            kernels[cur].append(line.strip())
    mangled = [k[1] for k in kernels]
    names = dict(zip(mangled, subprocess.run(["c++filt"], input="\n".join(mangled), capture_output=True, text=True).stdout.splitlines()))
    for family, sources in FAMILIES.items():
        out = [f"# SASS summary of {LIB.name}, {family} kernels -- `cuobjdump -sass` filtered by scripts/sass_summary.py", ""]
        for (s, fn), body in kernels.items():
            if s not in sources:
                continue
            count, shown = collections.Counter(), []
            for ins in body:
                m = re.match(r"/\*[0-9a-f]+\*/\s+(?:@!?U?P\w+\s+)?([A-Za-z0-9_.]+)", ins)
                if not m:
                    continue
                op, base = m.group(1), m.group(1).split(".")[0]
                if base not in COUNT:
                    continue
                count[op if base in SHOW else base] += 1
                if base in SHOW and sum(1 for x in shown if base in x) < 3:
                    shown.append("        " + re.sub(r"\s*/\* 0x[0-9a-f]+ \*/", "", ins))
            out.append(f"{s}.cu: {names.get(fn, fn)}")
            out.append(f"    {len(body)} instructions:  " + "  ".join(f"{k} x{v}" for k, v in sorted(count.items())))
            out.extend(shown)
            out.append("")
        path = ROOT / "profiles" / f"r02_sass_{family}.txt"
        path.write_text("\n".join(out))
        print(path.relative_to(ROOT), len(out), "lines")


if __name__ == "__main__":
    main()
