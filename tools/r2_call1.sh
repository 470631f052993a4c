#!/bin/bash
# round 2, call 1: GPU tests, full-size parity measurement, bench baseline
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/r2c1_tests.log 2>&1 ) 2>&1 | grep real; tail -3 gpurun_out/r2c1_tests.log | cut -c1-200
timeout 600 python tests/parity_full.py gpurun_out/r2c1_parity.json > gpurun_out/r2c1_parity.log 2>&1; tail -5 gpurun_out/r2c1_parity.log | cut -c1-300
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r2c1_bench.json 2> gpurun_out/r2c1_bench.err; tail -3 gpurun_out/r2c1_bench.err | cut -c1-300
python - <<'PY'
import json
try:
    d=json.load(open('gpurun_out/r2c1_bench.json'))
    print('value', d['value'], 'e2e', d['e2e']['value'], 'ms/step', d['ms_per_step'], 'frac', d['roofline']['frac'], 'launches', d['gpu_launches'], d['clocks'])
    print('kern', json.dumps(dict(list(d['roofline']['per_kernel_ms_per_step'].items())[:10])))
    print('cpu', d.get('cpu_baseline',{}).get('value'))
    for k,v in (d.get('other_workloads') or {}).items(): print(k, {a:b for a,b in v.items() if a!='api'})
except Exception as e: print('ERR', e)
PY
