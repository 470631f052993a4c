#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_pipeline_gpu.py tests/test_masks_gpu.py tests/test_det_gpu.py tests/test_isnet_gpu.py -x -q > gpurun_out/r2c7_tests.log 2>&1; tail -5 gpurun_out/r2c7_tests.log | cut -c1-300
timeout 600 python tools/kb_phases.py gpurun_out/r2c7_kb_phases.json 2>&1 | tail -40
