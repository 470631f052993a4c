#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_pipeline_gpu.py -x -q -m gpu > gpurun_out/t41.log 2>&1; tail -3 gpurun_out/t41.log | cut -c1-300
timeout 600 python bench.py --steps 3 --warmup 3 --no-other --no-cpu-baseline > gpurun_out/bench41.json 2> gpurun_out/bench41.err; tail -3 gpurun_out/bench41.err | cut -c1-300
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench41.json'))
print('value', d['value'], 'e2e', d['e2e']['value'], 'ms/step', d['ms_per_step'], 'frac', d['roofline']['frac'])
print('kern', json.dumps(d['roofline']['per_kernel_ms_per_step']))
PY
