#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_det_gpu.py tests/test_conv_gpu.py -x -q -m gpu > gpurun_out/t16.log 2>&1; tail -3 gpurun_out/t16.log
timeout 300 python tools/dw_bench.py 2>&1 | tee gpurun_out/dw16.log
timeout 600 python tools/det_profile.py 32 gpurun_out/det_profile7.json 2>&1 | tail -11
