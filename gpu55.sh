#!/bin/bash
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_zoe_dpt_gpu.py -x -q -m gpu -k "attention and tc" --durations=4 > gpurun_out/t55.log 2>&1; tail -12 gpurun_out/t55.log | cut -c1-200
echo "== tc"; timeout 400 python tools/zoe_bench.py 16 gpurun_out/zoe55_tc.json 2>&1 | grep -v Warn | tail -13
