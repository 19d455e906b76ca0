#!/bin/bash
# closing pass: tests after the preprocess store patch, the other BASELINE configs, the rasterizer-only workload, the default bench
O=gpurun_out; mkdir -p $O
timeout 400 python -m pytest tests/test_raster_gpu.py tests/test_decoder_gpu.py tests/test_pipeline_gpu.py -q -m gpu --timeout 120 2>&1 | tail -4
for c in 2 3 4; do
  timeout 400 python bench.py --config $c --steps 10 --warmup 3 > $O/r02_bench_config$c.json 2> $O/r02_bench_config$c.err; echo "config $c rc=$?"
done
timeout 300 python bench.py --workload splat --steps 50 --warmup 3 > $O/r02_bench_splat_n1.json 2> $O/r02_bench_splat.err; echo "splat rc=$?"
timeout 400 python bench.py --steps 20 --warmup 3 > $O/r02_bench_full_n1.json 2> $O/r02_bench_full.err; echo "full rc=$?"
python - <<'PY'
import json
for n in ("config2", "config3", "config4", "splat_n1", "full_n1"):
    try:
        d = json.loads(open(f"gpurun_out/r02_bench_{n}.json").read().strip().splitlines()[-1])
        st = {k: round(v["ms"], 3) for k, v in d.get("stages", {}).items() if isinstance(v, dict) and "ms" in v}
        print(n, {k: d[k] for k in ("value", "ms_per_step")}, "e2e", round(d["e2e"]["value"], 2), d["config"]["workload"][:60], st)
    except Exception as e:
        print(n, "FAILED", e)
PY
