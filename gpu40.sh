#!/bin/bash
mkdir -p gpurun_out
nproc; free -g | head -2
( time timeout 600 python bench.py --steps 3 --warmup 3 > gpurun_out/bench40.json 2> gpurun_out/bench40.err ) 2>&1 | grep real; tail -3 gpurun_out/bench40.err | cut -c1-300
python - <<'PY'
import json
try:
    d=json.load(open('gpurun_out/bench40.json'))
    print('value', d['value'], 'e2e', d['e2e']['value'], 'ms/step', d['ms_per_step'], 'frac', d['roofline']['frac'], 'launches', d['gpu_launches'])
    print('cpu', d.get('cpu_baseline'))
    print('other', json.dumps(d.get('other_workloads'))[:1500])
    print('kern', json.dumps(d['roofline']['per_kernel_ms_per_step']))
except Exception as e: print('ERR', e)
PY
( time timeout 900 python bench.py --depth zoe --steps 2 --warmup 3 --no-other --no-cpu-baseline > gpurun_out/bench40_zoe.json 2> gpurun_out/bench40_zoe.err ) 2>&1 | grep real; tail -3 gpurun_out/bench40_zoe.err | cut -c1-300
python - <<'PY'
import json
try:
    d=json.load(open('gpurun_out/bench40_zoe.json'))
    print('ZOE value', d['value'], 'e2e', d['e2e']['value'], 'ms/step', d['ms_per_step'], 'frac', d['roofline']['frac'])
    print('kern', json.dumps(d['roofline']['per_kernel_ms_per_step']))
except Exception as e: print('ERR', e)
PY
timeout 600 python -m pytest tests/test_zoe_dpt_gpu.py tests/test_zoe_gpu.py tests/test_pipeline_gpu.py -x -q -m gpu > gpurun_out/t40.log 2>&1; tail -3 gpurun_out/t40.log | cut -c1-300
