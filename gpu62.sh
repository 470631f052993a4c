#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kb_gpu.py tests/test_pipeline_gpu.py tests/test_det_gpu.py tests/test_leres_gpu.py tests/test_isnet_gpu.py tests/test_zoe_dpt_gpu.py -x -q -m gpu > gpurun_out/t62.log 2>&1; tail -3 gpurun_out/t62.log | cut -c1-300
timeout 600 python bench.py --steps 3 --warmup 3 --no-other --no-cpu-baseline > gpurun_out/bench62.json 2> gpurun_out/bench62.err; tail -3 gpurun_out/bench62.err | cut -c1-300
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench62.json'))
print('value', d['value'], 'e2e', d['e2e']['value'], 'ms/step', d['ms_per_step'], 'frac', d['roofline']['frac'])
print('kern', json.dumps(dict(list(d['roofline']['per_kernel_ms_per_step'].items())[:12])))
PY
