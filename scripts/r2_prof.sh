#!/bin/bash
# Round-2 profile pass: launch list of the timed step + ncu --set full captures of the dominant kernels (shares, not absolutes).
set -u
O=gpurun_out
mkdir -p $O
NCU_LIST="ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv"
FULL="ncu --set full --clock-control none"
# reports are reduced to their raw-metric CSV on the box (gpurun_out/ may carry at most 64 MiB back)
reduce() { ncu -i $O/$1.ncu-rep --page raw --csv > $O/$1.raw.csv 2>/dev/null; rm -f $O/$1.ncu-rep; }
timeout 400 $NCU_LIST --log-file $O/r02_launches_full.csv python bench.py --steps 1 --warmup 3 --profiler-range eager > $O/r02_ncu_full.log 2>&1; echo "list full rc=$?"
# convolutions named by the round-1 verdict: 7x7 128->256 @ 256^2 (N=8 context views) and 3x3 512->512 @ 64^2; plus the VAE's 3x3 128->128 @ 256^2
timeout 200 $FULL -k regex:k_conv -c 3 -o $O/r02_prof_conv_7x7 python scripts/conv_one.py 8 128 256 256 256 7 1 3 fwd 1 > $O/r02_conv1.log 2>&1; echo "conv 7x7 rc=$?"
reduce r02_prof_conv_7x7
timeout 200 $FULL -k regex:k_conv -c 3 -o $O/r02_prof_conv_3x3_512 python scripts/conv_one.py 4 512 64 64 512 3 1 1 fwd 1 > $O/r02_conv2.log 2>&1; echo "conv 3x3 512 rc=$?"
reduce r02_prof_conv_3x3_512
timeout 200 $FULL -k regex:k_conv -c 3 -o $O/r02_prof_conv_3x3_128 python scripts/conv_one.py 4 128 256 256 128 3 1 1 fwd 1 > $O/r02_conv3.log 2>&1; echo "conv 3x3 128 rc=$?"
reduce r02_prof_conv_3x3_128
timeout 200 $FULL -k regex:k_conv -c 3 -o $O/r02_prof_conv_w python scripts/conv_one.py 4 128 256 256 128 3 1 1 wgrad 1 > $O/r02_conv4.log 2>&1; echo "conv wgrad rc=$?"
reduce r02_prof_conv_w
timeout 200 $FULL -k regex:k_conv -c 3 -o $O/r02_prof_conv_d python scripts/conv_one.py 4 128 256 256 128 3 1 1 dgrad 1 > $O/r02_conv5.log 2>&1; echo "conv dgrad rc=$?"
reduce r02_prof_conv_d
# in-step captures
timeout 400 ncu --profile-from-start off --set full --clock-control none -k regex:k_gemm_tf32 -c 4 -o $O/r02_prof_gemm_step python bench.py --steps 1 --warmup 3 --profiler-range eager > $O/r02_ncu_gemm.log 2>&1; echo "gemm rc=$?"
reduce r02_prof_gemm_step
timeout 400 ncu --profile-from-start off --set full --clock-control none -k 'regex:k_preprocess|k_scatter_keys|k_tile_sort|k_blend' -c 6 -o $O/r02_prof_raster_step python bench.py --steps 1 --warmup 3 --profiler-range eager > $O/r02_ncu_raster.log 2>&1; echo "raster rc=$?"
reduce r02_prof_raster_step
timeout 400 ncu --profile-from-start off --set full --clock-control none -k 'regex:k_fmha' -c 4 -o $O/r02_prof_fmha_step python bench.py --steps 1 --warmup 3 --profiler-range eager > $O/r02_ncu_fmha.log 2>&1; echo "fmha rc=$?"
reduce r02_prof_fmha_step
ls -la $O/*.raw.csv $O/r02_launches_full.csv | tail -12
