#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_bokeh_gpu.py tests/test_pipeline_gpu.py tests/test_kb_gpu.py -x -q -m gpu > gpurun_out/t20.log 2>&1; tail -25 gpurun_out/t20.log
