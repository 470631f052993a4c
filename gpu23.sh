#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_zoe_dpt_gpu.py -x -q -m gpu -k "pipeline or end_to_end" > gpurun_out/t23.log 2>&1; tail -5 gpurun_out/t23.log | cut -c1-300
timeout 600 python tools/zoe_bench.py 16 gpurun_out/zoe_bench.json 2>&1 | tail -14
