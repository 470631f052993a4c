#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_det_gpu.py tests/test_pipeline_gpu.py -x -q -m gpu > gpurun_out/t29.log 2>&1; tail -5 gpurun_out/t29.log | cut -c1-300
timeout 600 python tools/det_profile.py 32 gpurun_out/det_profile9.json 2>&1 | tail -11
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/bench11.json 2> gpurun_out/bench11.err; tail -3 gpurun_out/bench11.err; python - <<'PY'
import json
d=json.load(open('gpurun_out/bench11.json'))
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'])
PY
