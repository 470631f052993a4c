#!/bin/bash
# round 2, call 34: row-blocked resample kernel A/B
mkdir -p gpurun_out
echo "== tests"; timeout 900 python -m pytest tests/test_det_gpu.py tests/test_leres_gpu.py tests/test_isnet_gpu.py tests/test_zoe_dpt_gpu.py tests/test_refine_gpu.py -q -m gpu -x 2>&1 | tail -3 | cut -c1-300
for m in 1 0 1 0; do
  CSB_RESAMPLE_ROWS=$m timeout 600 python bench.py --steps 10 --warmup 3 --no-other --no-cpu-baseline --stages depth > gpurun_out/r2c34_bench_$m.json 2> gpurun_out/r2c34_bench_$m.err; tail -2 gpurun_out/r2c34_bench_$m.err | cut -c1-300
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/r2c34_bench_$m.json'))
    pk=d['roofline']['per_kernel_ms_per_step']
    print('ROWS=$m leres depth-only ms/step', round(d['ms_per_step'],2), d['clocks']['sm_mhz'], 'k_resample', pk.get('k_resample'))
except Exception as e: print('ERR', e)
PY
done
for m in 1 0; do
  CSB_RESAMPLE_ROWS=$m timeout 600 python bench.py --depth zoe --steps 3 --warmup 3 --no-other --no-cpu-baseline --stages depth > gpurun_out/r2c34_zoe_$m.json 2> gpurun_out/r2c34_zoe_$m.err; tail -2 gpurun_out/r2c34_zoe_$m.err | cut -c1-300
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/r2c34_zoe_$m.json'))
    pk=d['roofline']['per_kernel_ms_per_step']
    print('ROWS=$m zoe depth-only ms/step', round(d['ms_per_step'],2), d['clocks']['sm_mhz'], 'k_resample', pk.get('k_resample'))
except Exception as e: print('ERR', e)
PY
done
