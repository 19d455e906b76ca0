#!/bin/bash
# after the elect.sync issue-path change in GEMM / conv: smoke, whole suite, bench, conv + gemm micro-benchmarks
mkdir -p gpurun_out
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
timeout 600 python -m pytest tests -q -m gpu --timeout 120 2>&1 | tail -6 | tee gpurun_out/r2_tests10.log
timeout 400 python bench.py --steps 20 --warmup 3 > gpurun_out/r2_bench10.json 2> gpurun_out/r2_bench10.err; python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2_bench10.json").read().strip().splitlines()[-1])
print({k: d[k] for k in ("value", "ms_per_step", "gpu_launches")}, d["e2e"]["value"], d.get("gpu_baseline", {}).get("value"))
print({k: (round(v["ms"], 2), round(v["achieved"])) for k, v in d["roofline"]["in_step_aggregate"].items() if isinstance(v, dict)})
print({k: round(v["achieved"]) for k, v in d["roofline"]["instances"].items()})
PY
timeout 200 python scripts/conv_bench.py 2>&1 | tail -25
