#!/bin/bash
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests -x -q -m gpu > gpurun_out/r2c9_tests.log 2>&1 ) 2>&1 | grep real; tail -4 gpurun_out/r2c9_tests.log | cut -c1-300
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r2c9_bench.json 2> gpurun_out/r2c9_bench.err; tail -3 gpurun_out/r2c9_bench.err | cut -c1-400
python - <<'PY'
import json
try:
    d=json.load(open('gpurun_out/r2c9_bench.json'))
    print('value', d['value'], 'e2e', d['e2e']['value'], 'ms/step', d['ms_per_step'], 'frac', d['roofline']['frac'], 'launches', d['gpu_launches'], d['clocks'])
    print('kern', json.dumps(dict(list(d['roofline']['per_kernel_ms_per_step'].items())[:12])))
    for k,v in (d.get('other_workloads') or {}).items(): print(k, json.dumps({a:b for a,b in v.items() if a!='api'})[:500])
except Exception as e: print('ERR', e)
PY
CSB_PROFILE_DETAIL=1 timeout 600 python tools/layer_profile.py 32 > gpurun_out/r2c9_layers.md 2>&1; tail -5 gpurun_out/r2c9_layers.md | cut -c1-200
