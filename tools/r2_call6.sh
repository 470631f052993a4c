#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/r2c6_tests.log 2>&1; tail -3 gpurun_out/r2c6_tests.log | cut -c1-300
timeout 600 python tools/kb_profile.py leres gpurun_out/r2c6_kb_profile.json > gpurun_out/r2c6_kb.log 2>&1; tail -60 gpurun_out/r2c6_kb.log
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2c6_bench.json 2> gpurun_out/r2c6_bench.err; tail -3 gpurun_out/r2c6_bench.err | cut -c1-300
python - <<'PY'
import json
try:
    d=json.load(open('gpurun_out/r2c6_bench.json'))
    print('value', d['value'], 'e2e', d['e2e']['value'], 'ms/step', d['ms_per_step'], 'frac', d['roofline']['frac'], 'launches', d['gpu_launches'], d['clocks'])
    print('kern', json.dumps(dict(list(d['roofline']['per_kernel_ms_per_step'].items())[:12])))
    for k,v in (d.get('other_workloads') or {}).items(): print(k, {a:b for a,b in v.items() if a!='api'})
except Exception as e: print('ERR', e)
PY
