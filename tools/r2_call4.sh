#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_dwconv_halo -s 1 -c 1 -f -o gpurun_out/r2c4_dw python tools/dw_stats_one.py 32 64 64 512 > gpurun_out/r2c4_ncu.log 2>&1; tail -2 gpurun_out/r2c4_ncu.log
