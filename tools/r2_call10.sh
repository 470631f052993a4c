#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/isnet_profile.py 20 720 gpurun_out/r2c10_isnet.json > gpurun_out/r2c10_isnet.log 2>&1; cat gpurun_out/r2c10_isnet.log | cut -c1-200
timeout 600 python tools/kb_hostprof.py > gpurun_out/r2c10_hostprof.log 2>&1; head -90 gpurun_out/r2c10_hostprof.log | cut -c1-220
