#!/bin/bash
# round 2, call 23: e2e with the H2D split per detector sub-batch (CSB_E2E_DET_SUB) + pipeline tests
mkdir -p gpurun_out
echo "== tests"; timeout 900 python -m pytest tests/test_pipeline_gpu.py tests/test_kb_gpu.py -q -m gpu -x 2>&1 | tail -4 | cut -c1-300
for sub in 32 16 8 32 16 8; do
  CSB_E2E_DET_SUB=$sub timeout 600 python bench.py --steps 10 --warmup 3 --no-other --no-cpu-baseline > gpurun_out/r2c23_bench_$sub.json 2> gpurun_out/r2c23_bench_$sub.err
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/r2c23_bench_$sub.json'))
    print('sub=$sub value', round(d['value'],1), 'ms/step', round(d['ms_per_step'],2), ' e2e', round(d['e2e']['value'],1), 'e2e ms', round(320*1e3/d['e2e']['value']/10,2), d['clocks']['sm_mhz'])
except Exception as e: print('ERR', e)
PY
done
