#!/bin/bash
# round 2, call 22: chunked top-k test (det_size > 1024) + stage ablation of the step (what the warp stage costs in the timed step)
mkdir -p gpurun_out
echo "== det tests"; timeout 900 python -m pytest tests/test_det_gpu.py -q -m gpu -x 2>&1 | tail -4 | cut -c1-300
for st in seg,depth,warp seg,depth warp seg depth; do
  timeout 600 python bench.py --steps 10 --warmup 3 --no-other --no-cpu-baseline --stages $st > gpurun_out/r2c22_bench.json 2> gpurun_out/r2c22_bench.err
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/r2c22_bench.json'))
    print('$st  ms/step', round(d['ms_per_step'],2), 'e2e ms', round(320*1e3/d['e2e']['value']/10,2), d['clocks']['sm_mhz'], 'launches', d['gpu_launches']//10)
except Exception as e: print('ERR', e)
PY
done
