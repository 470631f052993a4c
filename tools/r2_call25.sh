#!/bin/bash
# round 2, call 25: full GPU suite (net-io kernels, Lanczos tail, chunked top-k, hint waits), full bench with other workloads
mkdir -p gpurun_out
echo "== all gpu tests"; ( time timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/r2c25_tests.log 2>&1 ) 2>&1 | grep real; tail -6 gpurun_out/r2c25_tests.log | cut -c1-300
echo "== bench"
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2c25_bench.json 2> gpurun_out/r2c25_bench.err; tail -3 gpurun_out/r2c25_bench.err | cut -c1-400
python - <<'PY'
import json
try:
    d=json.load(open('gpurun_out/r2c25_bench.json'))
    print('value', d['value'], 'e2e', d['e2e']['value'], 'ms/step', d['ms_per_step'], 'frac', d['roofline']['frac'], 'launches', d['gpu_launches'], d['clocks'])
    print('cpu', d.get('cpu_baseline',{}).get('value'), 'parity', json.dumps(d.get('parity'))[:600])
    for k,v in (d.get('other_workloads') or {}).items(): print(k, json.dumps({a:b for a,b in v.items() if a not in('api',)})[:700])
except Exception as e: print('ERR', e)
PY
