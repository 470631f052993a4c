#!/bin/bash
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_zoe_gpu.py tests/test_zoe_dpt_gpu.py -x -q -m gpu -k "not mma_sync" > gpurun_out/t71.log 2>&1; tail -3 gpurun_out/t71.log | cut -c1-200
timeout 400 python tools/zoe_bench.py 16 gpurun_out/zoe71.json 2>&1 | grep -v Warn | tail -12
