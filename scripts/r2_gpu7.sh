#!/bin/bash
# fmha kernels + whole-suite regression + bench after wiring the attention cores
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_fmha_gpu.py -q -m gpu -x 2>&1 | tail -25
timeout 900 python -m pytest tests -q -m gpu -x --deselect tests/test_fmha_gpu.py 2>&1 | tail -8
timeout 400 python bench.py --steps 10 --warmup 3 > gpurun_out/r2_bench7.json 2> gpurun_out/r2_bench7.err; tail -c 1800 gpurun_out/r2_bench7.json; tail -3 gpurun_out/r2_bench7.err
