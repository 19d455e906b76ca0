#!/bin/bash
# closing validation of round 2: smoke, the whole GPU suite, the default bench line and the forward-only stress line
O=gpurun_out; mkdir -p $O
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -4
timeout 700 python -m pytest tests -q -m gpu --timeout 120 2>&1 | tail -4 | tee $O/r02_final_tests.log
timeout 500 python bench.py --steps 20 --warmup 3 > $O/r02_bench_full_n1.json 2> $O/r02_bench_full.err; echo "full rc=$?"
timeout 300 python bench.py --config 4 --steps 20 --warmup 3 > $O/r02_bench_config4.json 2> $O/r02_bench_config4.err; echo "config4 rc=$?"
timeout 200 python bench.py --impl reference --steps 1 --warmup 0 > $O/r02_bench_reference.json 2> $O/r02_bench_reference.err; echo "reference rc=$?"
python - <<'PY'
import json
for n in ("full_n1", "config4", "reference"):
    try:
        d = json.loads(open(f"gpurun_out/r02_bench_{n}.json").read().strip().splitlines()[-1])
        print(n, {k: d.get(k) for k in ("value", "ms_per_step", "gpu_launches")}, "e2e", d.get("e2e", {}).get("value"), d.get("roofline", {}) and (d["roofline"].get("kernel", "")[:40], d["roofline"].get("frac"), d["roofline"].get("traffic")))
    except Exception as e:
        print(n, "FAILED", e)
PY
