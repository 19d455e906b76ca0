#!/bin/bash
# GEMM cta_group::2 A/B on the whole step (tests + bench), and the config 2 / 3 lines with their corrected labels
O=gpurun_out; mkdir -p $O
LS_GEMM_2CTA=1 timeout 300 python -m pytest tests/test_gemm_gpu.py -q -m gpu --timeout 120 2>&1 | tail -3
LS_GEMM_2CTA=1 timeout 400 python bench.py --steps 20 --warmup 3 > $O/r02_bench_full_gemm2cta.json 2> $O/r02_bench_full_gemm2cta.err; echo "2cta rc=$?"
timeout 400 python bench.py --steps 20 --warmup 3 > $O/r02_bench_full_n1b.json 2> $O/r02_bench_full_n1b.err; echo "1cta rc=$?"
for c in 2 3; do
  timeout 400 python bench.py --config $c --steps 10 --warmup 3 > $O/r02_bench_config$c.json 2> $O/r02_bench_config$c.err; echo "config $c rc=$?"
done
python - <<'PY'
import json
for n in ("full_gemm2cta", "full_n1b", "config2", "config3"):
    try:
        d = json.loads(open(f"gpurun_out/r02_bench_{n}.json").read().strip().splitlines()[-1])
        agg = {k: (round(v["ms"], 2), round(v["achieved"])) for k, v in d["roofline"]["in_step_aggregate"].items() if isinstance(v, dict)}
        print(n, {k: round(d[k], 3) for k in ("value", "ms_per_step")}, d["config"]["workload"][:50], agg)
    except Exception as e:
        print(n, "FAILED", e)
PY
