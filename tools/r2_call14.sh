#!/bin/bash
mkdir -p gpurun_out
echo "== halo tests"; timeout 600 python -m pytest tests/test_halo_gpu.py -q 2>&1 | tail -5 | cut -c1-250
echo "== halo bench"; timeout 300 python tools/halo_bench.py gpurun_out/r2c14_halo_bench.json 2>&1 | tail -24
