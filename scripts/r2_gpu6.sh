#!/bin/bash
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests -q -m gpu > $O/r2_tests6.log 2>&1; tail -25 $O/r2_tests6.log
timeout 900 python bench.py --steps 10 --warmup 3 > $O/r2_bench5.json 2> $O/r2_bench5.err; python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/r2_bench5.json").read().strip().splitlines()[-1])
    print({k: d[k] for k in ("value", "ms_per_step", "gpu_launches")}, d["e2e"]["value"], d.get("gpu_baseline", {}).get("value"), {k: round(v["ms"], 4) for k, v in d["stages"].items() if "ms" in v})
except Exception as e:
    print("bench parse failed", e); print(open("gpurun_out/r2_bench5.err").read()[-3000:])
PY
