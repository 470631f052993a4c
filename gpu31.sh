#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_conv_gpu.py -x -q -m gpu > gpurun_out/t31a.log 2>&1; tail -4 gpurun_out/t31a.log | cut -c1-300
CSB_TMA_STORE=2 timeout 900 python -m pytest tests/test_conv_gpu.py -x -q -m gpu > gpurun_out/t31b.log 2>&1; echo "mode2:"; tail -2 gpurun_out/t31b.log | cut -c1-300
timeout 1200 python -m pytest tests/test_det_gpu.py tests/test_leres_gpu.py tests/test_isnet_gpu.py -x -q -m gpu > gpurun_out/t31c.log 2>&1; tail -3 gpurun_out/t31c.log | cut -c1-300
timeout 600 python tools/layer_profile.py 32 gpurun_out/layer_profile2.json 2>&1 | grep -E "^==|512->2048|128->512|256->1024|2048->512|256->256 k3x3 s1 d1 g1 act2|256->256 k1x1" | head -16
