#!/bin/bash
O=gpurun_out; mkdir -p $O
LS_CONV_2CTA=0 timeout 300 python -m pytest tests/test_conv_gpu.py -q -m gpu -x > $O/r2_conv_tests_1cta.log 2>&1; tail -4 $O/r2_conv_tests_1cta.log
LS_CONV_2CTA=2 timeout 300 python -m pytest tests/test_conv_gpu.py -q -m gpu > $O/r2_conv_tests_2cta.log 2>&1; tail -25 $O/r2_conv_tests_2cta.log
LS_CONV_2CTA=0 timeout 280 python scripts/conv_bench.py > $O/r2_conv_bench_1cta.log 2>&1; cat $O/r2_conv_bench_1cta.log
timeout 280 python scripts/conv_bench.py > $O/r2_conv_bench_2cta.log 2>&1; cat $O/r2_conv_bench_2cta.log
