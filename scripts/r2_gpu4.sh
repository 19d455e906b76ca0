#!/bin/bash
# round-2 GPU batch 4: rasterizer rework (sorted queues, compaction, v4 reds, padded gradients), fused GEMM epilogues, step bench
O=gpurun_out; mkdir -p $O
timeout 600 python -m pytest tests/test_raster_gpu.py tests/test_decoder_gpu.py -q -m gpu -x > $O/r2_raster_tests2.log 2>&1; tail -15 $O/r2_raster_tests2.log
timeout 600 python -m pytest tests -q -m gpu --deselect tests/test_raster_gpu.py --deselect tests/test_decoder_gpu.py > $O/r2_gpu_tests2.log 2>&1; tail -15 $O/r2_gpu_tests2.log
timeout 300 python bench.py --workload splat --steps 30 --warmup 3 > $O/r2_splat1.json 2> $O/r2_splat1.err; python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/r2_splat1.json").read().strip().splitlines()[-1])
    print("splat", d["value"], d["ms_per_step"], {k: round(v["ms"], 4) for k, v in d["stages"].items()})
except Exception as e:
    print("splat parse failed", e); print(open("gpurun_out/r2_splat1.err").read()[-2000:])
PY
timeout 900 python bench.py --steps 10 --warmup 3 > $O/r2_bench3.json 2> $O/r2_bench3.err; python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/r2_bench3.json").read().strip().splitlines()[-1])
    print({k: d[k] for k in ("value", "ms_per_step", "gpu_launches")}, d["e2e"], d.get("gpu_baseline"), {k: round(v["ms"], 4) for k, v in d["stages"].items() if "ms" in v})
except Exception as e:
    print("bench parse failed", e); print(open("gpurun_out/r2_bench3.err").read()[-3000:])
PY
