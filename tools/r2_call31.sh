#!/bin/bash
# round 2, call 31: half-warp-per-pixel LayerNorm (C <= 128): test + step A/B
mkdir -p gpurun_out
echo "== tests"; timeout 600 python -m pytest tests/test_det_gpu.py tests/test_leres_gpu.py -q -m gpu -x -k "elementwise or network_forward or lanczos" 2>&1 | tail -3 | cut -c1-300
for m in 1 0 1 0; do
  CSB_LN_H16=$m timeout 600 python bench.py --steps 10 --warmup 3 --no-other --no-cpu-baseline > gpurun_out/r2c31_bench_$m.json 2> gpurun_out/r2c31_bench_$m.err; tail -2 gpurun_out/r2c31_bench_$m.err | cut -c1-300
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/r2c31_bench_$m.json'))
    pk=d['roofline']['per_kernel_ms_per_step']
    print('LN_H16=$m value', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), 'ms/step', round(d['ms_per_step'],2), d['clocks']['sm_mhz'], 'k_layernorm', pk.get('k_layernorm'), 'k_resample', pk.get('k_resample'))
except Exception as e: print('ERR', e)
PY
done
