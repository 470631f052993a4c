#!/bin/bash
# round 2, call 15: measurement evidence -- ncu DRAM traffic + launch list of the bench step, per-layer profile, ISNet profile, full GPU tests + bench
mkdir -p gpurun_out
echo "== ncu traffic / launch list of the bench step"
( time timeout 700 ncu -c 1500 --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:k_ --csv \
    --log-file gpurun_out/r2c15_traffic.csv python bench.py --steps 1 --warmup 3 --no-other --no-cpu-baseline > gpurun_out/r2c15_ncu_bench.log 2>&1 ) 2>&1 | grep real
ls -la gpurun_out/r2c15_traffic.csv
python tools/ncu_traffic.py gpurun_out/r2c15_traffic.csv gpurun_out/r2c15_traffic.json --batch 32 --stages seg,depth,warp --depth leres 2>&1 | tail -14
gzip -f gpurun_out/r2c15_traffic.csv
echo "== layer profile"
timeout 600 python tools/layer_profile.py 32 gpurun_out/r2c15_layers.json > gpurun_out/r2c15_layers.md 2>&1; head -40 gpurun_out/r2c15_layers.md | cut -c1-160
echo "== isnet profile"
timeout 600 python tools/isnet_profile.py 20 720 gpurun_out/r2c15_isnet.json > gpurun_out/r2c15_isnet.log 2>&1; head -40 gpurun_out/r2c15_isnet.log | cut -c1-160
echo "== bench"
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r2c15_bench.json 2> gpurun_out/r2c15_bench.err; tail -3 gpurun_out/r2c15_bench.err | cut -c1-400
python - <<'PY'
import json
try:
    d=json.load(open('gpurun_out/r2c15_bench.json'))
    print('value', d['value'], 'e2e', d['e2e']['value'], 'ms/step', d['ms_per_step'], 'frac', d['roofline']['frac'], 'launches', d['gpu_launches'], d['clocks'])
    print('cpu', d.get('cpu_baseline'))
    for k,v in (d.get('other_workloads') or {}).items(): print(k, json.dumps({a:b for a,b in v.items() if a not in('api','per_kernel_ms_profiled')})[:500])
except Exception as e: print('ERR', e)
PY
