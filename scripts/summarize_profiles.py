#!/usr/bin/env python
"""Turn the scratch ncu outputs in gpurun_out/ into the tracked summaries under profiles/.

  python scripts/summarize_profiles.py r01 [launches.csv] [prof.ncu-rep]
"""
import collections
import csv
import json
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
launches = Path(sys.argv[2]) if len(sys.argv) > 2 else ROOT / "gpurun_out" / "launches.csv"
rep = Path(sys.argv[3]) if len(sys.argv) > 3 else ROOT / "gpurun_out" / "prof_blend.ncu-rep"
out = ROOT / "profiles"
out.mkdir(exist_ok=True)

if launches.exists():
    lines = [l for l in launches.read_text().splitlines(True) if not l.startswith("==")]
    rows = list(csv.DictReader(lines))
    agg = collections.OrderedDict()
    for x in rows:
        try:
            t = float(x["Metric Value"].replace(",", ""))
        except ValueError:
            continue
        if t != t:          # nan: a launch ncu could not time
            continue
        t = {"ns": t / 1e3, "us": t, "usecond": t, "ms": t * 1e3, "msecond": t * 1e3, "nsecond": t / 1e3}.get(x["Metric Unit"], t)
        a = agg.setdefault(x["Kernel Name"][:90], [0, 0.0, x["Grid Size"], x["Block Size"]])
        a[0] += 1
        a[1] += t
    tot = sum(a[1] for a in agg.values())
    import re
    ours = sum(a[1] for k, a in agg.items() if re.search(r"\bls[a-z]?::", k))
    md = [f"# {tag}: ncu launch list of the timed region of `python bench.py` (cudaProfilerStart/Stop range; gpu__time_duration.sum, --clock-control none)",
          "", f"{len(rows)} launches, {tot / 1e3:.2f} ms GPU time in total; our kernels (`ls::*`, `lsg::*`, `lsc::*`, `lsf::*`, `lsa::*`, `lse::*`, `lsh::*`, `lsn::*`) = {100 * ours / tot:.1f} % of it.",
          "Per-launch times under ncu are cold-cache and serialised: compare SHARES, not absolutes.", "",
          "| share | total µs | launches | avg µs | grid | block | kernel |", "|---:|---:|---:|---:|---|---|---|"]
    for k, (n, t, g, b) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:30]:
        md.append(f"| {100 * t / tot:.1f} % | {t:.1f} | {n} | {t / n:.1f} | {g} | {b} | `{k}` |")
    (out / f"{tag}_launches.md").write_text("\n".join(md) + "\n")
    print("wrote", out / f"{tag}_launches.md")

if rep.exists():
    # a .ncu-rep is read through ncu; a .csv is the `--page raw --csv` dump already made on the GPU box
    raw = rep.read_text() if rep.suffix == ".csv" else \
        subprocess.run(["ncu", "-i", str(rep), "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
            "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
            "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
            "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
            "inst_executed", "smsp__thread_inst_executed_per_inst_executed.ratio",
            "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_red.sum",
            "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
            "TPC.TriageCompute.sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
            "sm__inst_executed_pipe_tmem.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
            "sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum", "launch__grid_size",
            "launch__block_size", "smsp__cycles_active.avg"]
    md = [f"# {tag}: ncu --set full capture ({rep.name})", ""]
    traffic = {}
    to_bytes = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")]
        md += [f"## `{name}`", "", "| metric | value | unit |", "|---|---:|---|"]
        rd = wr = None
        for w in want:
            if w in hdr:
                i = hdr.index(w)
                md.append(f"| {w} | {r[i]} | {units[i]} |")
                if w == "dram__bytes_read.sum":
                    rd = float(r[i].replace(",", "")) * to_bytes.get(units[i], 1)
                if w == "dram__bytes_write.sum":
                    wr = float(r[i].replace(",", "")) * to_bytes.get(units[i], 1)
        md.append("")
        key = "blend_fwd" if "blend_fwd" in name else "blend_bwd" if "blend_bwd" in name else \
            "preprocess_bwd" if "preprocess_bwd" in name else "preprocess" if "preprocess" in name else \
            "sort" if "sort" in name else "scatter" if "scatter" in name else name
        if rd is not None and wr is not None:
            traffic[key] = {"dram_bytes_per_launch": rd + wr, "kernel": name}
    (out / f"{tag}_ncu_{rep.stem}.md").write_text("\n".join(md) + "\n")
    tj = out / f"{tag}_ncu_traffic.json"
    old = json.loads(tj.read_text()) if tj.exists() else {}
    old.update(traffic)
    tj.write_text(json.dumps(old, indent=1) + "\n")
    print("wrote", out / f"{tag}_ncu_{rep.stem}.md", tj)
