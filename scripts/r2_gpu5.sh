#!/bin/bash
O=gpurun_out; mkdir -p $O
timeout 300 python scripts/gemm_epi.py 2>&1 | tail -5
timeout 600 python -m pytest tests/test_gemm_gpu.py tests/test_raster_gpu.py -q -m gpu > $O/r2_tests5.log 2>&1; tail -6 $O/r2_tests5.log
LS_GEMM_2CTA=1 timeout 300 python scripts/gemm_epi.py 2>&1 | tail -5
timeout 900 python bench.py --steps 10 --warmup 3 > $O/r2_bench4.json 2> $O/r2_bench4.err; python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/r2_bench4.json").read().strip().splitlines()[-1])
    print({k: d[k] for k in ("value", "ms_per_step", "gpu_launches")}, d["e2e"]["value"], d.get("gpu_baseline", {}).get("value"), {k: round(v["ms"], 4) for k, v in d["stages"].items() if "ms" in v})
except Exception as e:
    print("bench parse failed", e); print(open("gpurun_out/r2_bench4.err").read()[-3000:])
PY
