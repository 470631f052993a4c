#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_zoe_dpt_gpu.py -x -q -m gpu -k "attention" > gpurun_out/t53.log 2>&1; tail -25 gpurun_out/t53.log | cut -c1-300
