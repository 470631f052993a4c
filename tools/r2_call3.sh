#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_dw_halo_gpu.py -x -q > gpurun_out/r2c3_dw_tests.log 2>&1; tail -5 gpurun_out/r2c3_dw_tests.log | cut -c1-300
timeout 300 python tools/dw_bench2.py 2>&1 | tail -2
