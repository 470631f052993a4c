#!/bin/bash
# round 2, call 41: final records with the final tree -- full GPU suite (default paths), detector / conv suites on the legacy switches, smoke, bench, --depth zoe
mkdir -p gpurun_out
echo "== all gpu tests"; ( time timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/r2c41_tests.log 2>&1 ) 2>&1 | grep real; tail -4 gpurun_out/r2c41_tests.log | cut -c1-300
echo "== legacy switches (two-launch MLP, polynomial GELU, per-tap only, no CTA pairs)"
CSB_FUSE_MLP=0 CSB_GELU_FORM=poly CSB_CTA_PAIR=0 timeout 900 python -m pytest tests/test_det_gpu.py tests/test_conv_gpu.py tests/test_parity_full_gpu.py -q -m gpu 2>&1 | tail -3 | cut -c1-300
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "== bench"
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2c41_bench.json 2> gpurun_out/r2c41_bench.err; tail -3 gpurun_out/r2c41_bench.err | cut -c1-400
python - <<'PY'
import json
try:
    d=json.load(open('gpurun_out/r2c41_bench.json'))
    r=d['roofline']
    print('value', d['value'], 'e2e', d['e2e']['value'], 'ms/step', d['ms_per_step'], 'frac', r['frac'], 'traffic/alg', r.get('traffic_over_algorithmic'), 'launches', d['gpu_launches'], d['clocks'])
    print('engine', json.dumps(r.get('conv_engine',{}).get('per_kernel')), r.get('conv_engine',{}).get('frac'))
    print('kern', json.dumps(dict(list(r['per_kernel_ms_per_step'].items())[:12])))
    print('cpu', d.get('cpu_baseline',{}).get('value'))
    for k,v in (d.get('other_workloads') or {}).items(): print(k, json.dumps({a:b for a,b in v.items() if a not in('api','per_kernel_ms_profiled')})[:420])
except Exception as e: print('ERR', e)
PY
echo "== bench --depth zoe"
timeout 900 python bench.py --depth zoe --steps 3 --warmup 3 --no-other --no-cpu-baseline > gpurun_out/r2c41_bench_zoe.json 2> gpurun_out/r2c41_bench_zoe.err; tail -2 gpurun_out/r2c41_bench_zoe.err | cut -c1-300
python - <<'PY'
import json
try:
    d=json.load(open('gpurun_out/r2c41_bench_zoe.json'))
    print('ZOE value', d['value'], 'e2e', d['e2e']['value'], 'ms/step', d['ms_per_step'], 'frac', d['roofline']['frac'], d['clocks'])
except Exception as e: print('ERR', e)
PY
