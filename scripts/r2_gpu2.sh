#!/bin/bash
# round-2 GPU batch 2: conv kernels after the epilogue / schedule rework, raster parity at full size, step launch list
O=gpurun_out; mkdir -p $O
timeout 300 python -m pytest tests/test_conv_gpu.py -q -m gpu > $O/r2_conv_tests_t.log 2>&1; tail -3 $O/r2_conv_tests_t.log
timeout 280 python scripts/conv_bench.py > $O/r2_conv_bench_t2.log 2>&1; cat $O/r2_conv_bench_t2.log
timeout 600 python -m pytest tests/test_raster_gpu.py -q -m gpu > $O/r2_raster_tests.log 2>&1; tail -30 $O/r2_raster_tests.log
timeout 600 python bench.py --steps 10 --warmup 3 > $O/r2_bench2.json 2> $O/r2_bench2.err; python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/r2_bench2.json").read().strip().splitlines()[-1])
    print({k: d[k] for k in ("value", "ms_per_step", "gpu_launches")}, d["e2e"], d["config"].get("eager_exact_ms_per_step"))
except Exception as e:
    print("bench parse failed", e); print(open("gpurun_out/r2_bench2.err").read()[-3000:])
PY
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/r2_launches_full.csv python bench.py --steps 1 --warmup 3 --profiler-range eager > $O/r2_ncu_full.log 2>&1; echo "list rc=$?"
python scripts/summarize_profiles.py r02_full_tmp $O/r2_launches_full.csv /nonexistent > /dev/null 2>&1; head -60 profiles/r02_full_tmp_launches.md; cp profiles/r02_full_tmp_launches.md $O/
