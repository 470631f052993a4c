#!/bin/bash
# round 2, call 29: final evidence -- full GPU suite, ncu DRAM traffic + launch list of the bench step (with k_mlp_tc), full bench
mkdir -p gpurun_out
echo "== all gpu tests"; ( time timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/r2c29_tests.log 2>&1 ) 2>&1 | grep real; tail -4 gpurun_out/r2c29_tests.log | cut -c1-300
echo "== ncu traffic / launch list of the bench step"
( time timeout 700 ncu -c 1500 --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:k_ --csv \
    --log-file gpurun_out/r2c29_traffic.csv python bench.py --steps 1 --warmup 3 --no-other --no-cpu-baseline > gpurun_out/r2c29_ncu_bench.log 2>&1 ) 2>&1 | grep real
python tools/ncu_traffic.py gpurun_out/r2c29_traffic.csv gpurun_out/r2c29_traffic.json --batch 32 --stages seg,depth,warp --depth leres --steps 2 2>&1 | tail -14
gzip -f gpurun_out/r2c29_traffic.csv
echo "== bench"
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2c29_bench.json 2> gpurun_out/r2c29_bench.err; tail -3 gpurun_out/r2c29_bench.err | cut -c1-400
python - <<'PY'
import json
try:
    d=json.load(open('gpurun_out/r2c29_bench.json'))
    print('value', d['value'], 'e2e', d['e2e']['value'], 'ms/step', d['ms_per_step'], 'frac', d['roofline']['frac'], 'launches', d['gpu_launches'], d['clocks'])
    print('engine', json.dumps(d['roofline'].get('conv_engine'))[:700])
    print('cpu', d.get('cpu_baseline',{}).get('value'))
    for k,v in (d.get('other_workloads') or {}).items(): print(k, json.dumps({a:b for a,b in v.items() if a not in('api','per_kernel_ms_profiled')})[:500])
except Exception as e: print('ERR', e)
PY
