#!/bin/bash
# fmha kernels first (bounded), then the new tests, then the whole suite + bench; every pytest call has per-test timeouts
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_fmha_gpu.py -q -m gpu -x --timeout 60 2>&1 | tail -15 | tee gpurun_out/r2_fmha9.log
if grep -q "failed\|Timeout\|error" gpurun_out/r2_fmha9.log; then echo "fmha failed: stopping"; exit 1; fi
timeout 600 python -m pytest tests -q -m gpu --timeout 120 2>&1 | tail -30 | tee gpurun_out/r2_tests9.log
timeout 400 python bench.py --steps 10 --warmup 3 > gpurun_out/r2_bench9.json 2> gpurun_out/r2_bench9.err; tail -c 2500 gpurun_out/r2_bench9.json; tail -3 gpurun_out/r2_bench9.err
