#!/bin/bash
mkdir -p gpurun_out
echo "== pair tests"; timeout 300 python -m pytest tests/test_conv_gpu.py -x -q -k "cta_pair" 2>&1 | tail -15 | cut -c1-250
echo "== conv tests, pair off"; CSB_CTA_PAIR=0 timeout 300 python -m pytest tests/test_conv_gpu.py -x -q 2>&1 | tail -5 | cut -c1-250
echo "== conv tests, default"; timeout 300 python -m pytest tests/test_conv_gpu.py -x -q 2>&1 | tail -5 | cut -c1-250
echo "== pair bench"; timeout 300 python tools/pair_bench.py gpurun_out/r2c11_pair_bench.json 2>&1 | tail -20
