#!/bin/bash
# Short closing pass: smoke, GPU tests, both bench workloads, launch lists (the --set full captures come from final_gpu.sh).
set -u
O=gpurun_out
mkdir -p $O
NCU_LIST="ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/final_smoke.log 2>&1; echo "smoke rc=$?"
timeout 600 python -m pytest tests -m gpu -q > $O/final_tests.log 2>&1; tail -1 $O/final_tests.log
timeout 600 python bench.py --steps 20 --warmup 3 > $O/final_full.json 2> $O/final_full.err; echo "full rc=$?"
timeout 300 python bench.py --workload splat --steps 50 --warmup 3 > $O/final_splat.json 2> $O/final_splat.err; echo "splat rc=$?"
timeout 600 $NCU_LIST --log-file $O/launches_full.csv python bench.py --steps 1 --warmup 3 --profiler-range eager > $O/ncu_full.log 2>&1; echo "list full rc=$?"
timeout 300 $NCU_LIST --log-file $O/launches_splat.csv python bench.py --workload splat --steps 2 --warmup 3 --profiler-range > $O/ncu_splat.log 2>&1; echo "list splat rc=$?"
