#!/bin/bash
mkdir -p gpurun_out
echo "== halo tests"; timeout 600 python -m pytest tests/test_halo_gpu.py -q -rs 2>&1 | tail -60 | cut -c1-250
echo "== conv tests"; timeout 300 python -m pytest tests/test_conv_gpu.py -x -q 2>&1 | tail -5 | cut -c1-250
