#!/bin/bash
# One gpurun call: smoke, GPU tests, benches, ncu launch lists and --set full captures.  Outputs under gpurun_out/.
# Numbers printed by runs under ncu are never bench values.
set -u
O=gpurun_out
mkdir -p $O
NCU_LIST="ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv"
NCU_FULL="ncu --profile-from-start off --set full --clock-control none --import-source on"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/final_smoke.log 2>&1; echo "smoke rc=$?"
timeout 600 python -m pytest tests -m gpu -q > $O/final_tests.log 2>&1; tail -2 $O/final_tests.log
timeout 600 python bench.py --steps 20 --warmup 3 > $O/final_full.json 2> $O/final_full.err; echo "full rc=$?"
timeout 300 python bench.py --workload splat --steps 50 --warmup 3 > $O/final_splat.json 2> $O/final_splat.err; echo "splat rc=$?"
timeout 600 $NCU_LIST --log-file $O/launches_full.csv python bench.py --steps 1 --warmup 3 --profiler-range eager > $O/ncu_full.log 2>&1; echo "list full rc=$?"
timeout 300 $NCU_LIST --log-file $O/launches_splat.csv python bench.py --workload splat --steps 2 --warmup 3 --profiler-range > $O/ncu_splat.log 2>&1; echo "list splat rc=$?"
timeout 600 $NCU_FULL -k regex:k_gemm_tf32 -c 4 -f -o $O/prof_gemm_step python bench.py --steps 1 --warmup 3 --profiler-range > $O/ncu_gemm.log 2>&1; echo "gemm rc=$?"
timeout 600 $NCU_FULL -k 'regex:k_preprocess|k_scatter_keys|k_tile_sort|k_blend' -c 6 -f -o $O/prof_raster_step python bench.py --workload splat --steps 1 --warmup 3 --profiler-range > $O/ncu_raster.log 2>&1; echo "raster rc=$?"
timeout 600 $NCU_FULL -k 'regex:k_gather|k_absorbed' -c 6 -f -o $O/prof_epipolar_step python bench.py --steps 1 --warmup 3 --profiler-range > $O/ncu_epi.log 2>&1; echo "epi rc=$?"
timeout 600 $NCU_FULL -k 'regex:k_gn_' -c 4 -f -o $O/prof_gn_step python bench.py --steps 1 --warmup 3 --profiler-range > $O/ncu_gn.log 2>&1; echo "gn rc=$?"
timeout 600 $NCU_FULL -k 'regex:k_ln_' -c 2 -f -o $O/prof_ln_step python bench.py --steps 1 --warmup 3 --profiler-range > $O/ncu_ln.log 2>&1; echo "ln rc=$?"
ls -la $O | tail -30
