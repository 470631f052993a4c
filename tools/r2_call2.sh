#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_parity_full_gpu.py -x -q -s > gpurun_out/r2c2_parity_tests.log 2>&1; tail -5 gpurun_out/r2c2_parity_tests.log | cut -c1-300
timeout 900 python -m pytest tests -q -m gpu -s --deselect tests/test_parity_full_gpu.py > gpurun_out/r2c2_tests_s.log 2>&1; tail -3 gpurun_out/r2c2_tests_s.log | cut -c1-300
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_dwconv_tile -s 4 -c 1 -f -o gpurun_out/r2c2_dw python tools/dw_stats_one.py 16 64 64 512 > gpurun_out/r2c2_ncu.log 2>&1; tail -2 gpurun_out/r2c2_ncu.log
