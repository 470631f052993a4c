#!/bin/bash
mkdir -p gpurun_out
( time timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench60.json 2> gpurun_out/bench60.err ) 2>&1 | grep real; tail -3 gpurun_out/bench60.err | cut -c1-300
python - <<'PY'
import json
try:
    d=json.load(open('gpurun_out/bench60.json'))
    print('value', d['value'], 'e2e', d['e2e']['value'], 'ms/step', d['ms_per_step'], 'frac', d['roofline']['frac'], 'launches', d['gpu_launches'], d['clocks'])
    print('cpu', d.get('cpu_baseline'))
    for k,v in (d.get('other_workloads') or {}).items(): print(k, {a:b for a,b in v.items() if a!='api'})
    print('kern', json.dumps(d['roofline']['per_kernel_ms_per_step']))
except Exception as e: print('ERR', e)
PY
( time timeout 900 python bench.py --depth zoe --steps 3 --warmup 3 --no-other --no-cpu-baseline > gpurun_out/bench60_zoe.json 2> gpurun_out/bench60_zoe.err ) 2>&1 | grep real; tail -3 gpurun_out/bench60_zoe.err | cut -c1-300
python - <<'PY'
import json
try:
    d=json.load(open('gpurun_out/bench60_zoe.json'))
    print('ZOE value', d['value'], 'e2e', d['e2e']['value'], 'ms/step', d['ms_per_step'], 'frac', d['roofline']['frac'])
    print('kern', json.dumps(d['roofline']['per_kernel_ms_per_step']))
except Exception as e: print('ERR', e)
PY
