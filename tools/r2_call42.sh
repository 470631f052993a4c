#!/bin/bash
# round 2, call 42: sanity of the final bench.py (median-of-three profile) -- default invocation as the driver runs it
mkdir -p gpurun_out
( time timeout 900 python bench.py > gpurun_out/r2c42_bench.json 2> gpurun_out/r2c42_bench.err ) 2>&1 | grep real; tail -2 gpurun_out/r2c42_bench.err | cut -c1-300
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2c42_bench.json'))
r=d['roofline']
print('value', d['value'], 'e2e', d['e2e']['value'], 'ms/step', d['ms_per_step'], 'steps', d['steps'], 'frac', r['frac'], 'traffic/alg', r.get('traffic_over_algorithmic'), d['clocks'])
print('engine frac', r['conv_engine']['frac'], 'cpu', d['cpu_baseline']['value'], 'parity', json.dumps(d['parity']['warped_frame_u8']))
print(sorted(d.keys()))
PY
