#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_dw_halo_gpu.py tests/test_det_gpu.py -x -q > gpurun_out/r2c5_tests.log 2>&1; tail -3 gpurun_out/r2c5_tests.log | cut -c1-300
timeout 300 python tools/dw_bench2.py 2>&1 | tail -2
