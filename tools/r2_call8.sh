#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/r2c8_tests.log 2>&1; tail -4 gpurun_out/r2c8_tests.log | cut -c1-300
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r2c8_bench.json 2> gpurun_out/r2c8_bench.err; tail -5 gpurun_out/r2c8_bench.err | cut -c1-400
python - <<'PY'
import json
try:
    d=json.load(open('gpurun_out/r2c8_bench.json'))
    print('value', d['value'], 'e2e', d['e2e']['value'], 'ms/step', d['ms_per_step'], 'frac', d['roofline']['frac'], 'launches', d['gpu_launches'], d['clocks'])
    print('kern', json.dumps(dict(list(d['roofline']['per_kernel_ms_per_step'].items())[:12])))
    print('cpu', d.get('cpu_baseline'))
    print('cpu512', d.get('cpu_baseline_512'))
    print('parity', json.dumps(d.get('parity'))[:1500])
    for k,v in (d.get('other_workloads') or {}).items(): print(k, json.dumps({a:b for a,b in v.items() if a!='api'})[:700])
except Exception as e: print('ERR', e)
PY
