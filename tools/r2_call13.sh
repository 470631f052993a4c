#!/bin/bash
mkdir -p gpurun_out
echo "== halo bench"; timeout 300 python tools/halo_bench.py gpurun_out/r2c13_halo_bench.json 2>&1 | tail -24
echo "== all gpu tests"; timeout 1200 python -m pytest tests -x -q -m gpu > gpurun_out/r2c13_tests.log 2>&1; tail -15 gpurun_out/r2c13_tests.log | cut -c1-300
echo "== bench"; timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2c13_bench.json 2> gpurun_out/r2c13_bench.err; tail -3 gpurun_out/r2c13_bench.err | cut -c1-400
python - <<'PY'
import json
try:
    d=json.load(open('gpurun_out/r2c13_bench.json'))
    print('value', d['value'], 'e2e', d['e2e']['value'], 'ms/step', d['ms_per_step'], 'frac', d['roofline']['frac'], 'launches', d['gpu_launches'], d['clocks'])
    print('kern', json.dumps(dict(list(d['roofline']['per_kernel_ms_per_step'].items())[:14])))
    for k,v in (d.get('other_workloads') or {}).items(): print(k, json.dumps({a:b for a,b in v.items() if a!='api'})[:600])
except Exception as e: print('ERR', e)
PY
