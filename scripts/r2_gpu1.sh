#!/bin/bash
# round-2 GPU batch 1: conv kernels in both orientations, conv bench, whole GPU suite, a short full-step bench
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_conv_gpu.py -q -m gpu > gpurun_out/r2_conv_tests_t.log 2>&1
LS_CONV_ORIENT=f timeout 300 python -m pytest tests/test_conv_gpu.py -q -m gpu > gpurun_out/r2_conv_tests_f.log 2>&1
timeout 280 python scripts/conv_bench.py > gpurun_out/r2_conv_bench_t.log 2>&1
timeout 900 python -m pytest tests -q -m gpu -x --deselect tests/test_conv_gpu.py > gpurun_out/r2_gpu_tests.log 2>&1
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r2_bench1.json 2> gpurun_out/r2_bench1.err
tail -5 gpurun_out/r2_conv_tests_t.log; tail -5 gpurun_out/r2_conv_tests_f.log; cat gpurun_out/r2_conv_bench_t.log; tail -15 gpurun_out/r2_gpu_tests.log; tail -c 3000 gpurun_out/r2_bench1.json; tail -5 gpurun_out/r2_bench1.err
