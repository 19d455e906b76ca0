#!/bin/bash
O=gpurun_out; mkdir -p $O
timeout 200 python scripts/gemm_layouts.py 2>&1 | tail -6
timeout 300 python -m pytest tests/test_conv_gpu.py -q -m gpu -k "8x8x0 or forward" 2>&1 | tail -3
ncu --set full --clock-control none --import-source on -k regex:k_conv_tf32 -s 1 -c 1 -f -o $O/r2_ncu_conv_w2 python scripts/conv_one.py 2 128 256 256 256 7 1 3 wgrad 2 > $O/ncu_w2.log 2>&1; tail -2 $O/ncu_w2.log
