#!/bin/bash
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests -x -q -m gpu > gpurun_out/t70.log 2>&1 ) 2>&1 | grep real; tail -3 gpurun_out/t70.log | cut -c1-200
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench70.json 2> gpurun_out/bench70.err; tail -3 gpurun_out/bench70.err | cut -c1-300
python - <<'PY'
import json
try:
    d=json.load(open('gpurun_out/bench70.json'))
    print('value', d['value'], 'e2e', d['e2e']['value'], 'ms/step', d['ms_per_step'], 'frac', d['roofline']['frac'], 'launches', d['gpu_launches'], d['clocks'])
    print('cpu', d.get('cpu_baseline',{}).get('value'))
    for k,v in (d.get('other_workloads') or {}).items(): print(k, {a:b for a,b in v.items() if a!='api'})
except Exception as e: print('ERR', e)
PY
timeout 900 python bench.py --depth zoe --steps 3 --warmup 3 --no-other --no-cpu-baseline > gpurun_out/bench70_zoe.json 2> gpurun_out/bench70_zoe.err; tail -3 gpurun_out/bench70_zoe.err | cut -c1-300
python - <<'PY'
import json
try:
    d=json.load(open('gpurun_out/bench70_zoe.json'))
    print('ZOE value', d['value'], 'e2e', d['e2e']['value'], 'ms/step', d['ms_per_step'], 'frac', d['roofline']['frac'])
    print('kern', json.dumps(dict(list(d['roofline']['per_kernel_ms_per_step'].items())[:8])))
except Exception as e: print('ERR', e)
PY
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
