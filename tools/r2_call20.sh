#!/bin/bash
# round 2, call 20: A-resident n-loop of k_conv_tc -- correctness (bit-identical to the per-tile pipeline) and speed
mkdir -p gpurun_out
echo "== tests"; timeout 900 python -m pytest tests/test_conv_gpu.py -q -m gpu -x 2>&1 | tail -6 | cut -c1-300
for m in 0 1; do
  echo "== gelu_check CSB_A_RES=$m"; CSB_A_RES=$m timeout 300 python tools/gelu_check.py 2>&1 | tail -3
  echo "== no act CSB_A_RES=$m"; ACT=none CSB_A_RES=$m timeout 300 python tools/gelu_check.py 2>&1 | tail -3
done
echo "== det tests"; timeout 900 python -m pytest tests/test_det_gpu.py tests/test_parity_full_gpu.py -q -m gpu -x 2>&1 | tail -4 | cut -c1-300
for m in 1 0 1 0; do
  CSB_A_RES=$m timeout 600 python bench.py --steps 10 --warmup 3 --no-other --no-cpu-baseline > gpurun_out/r2c20_bench_$m.json 2> gpurun_out/r2c20_bench_$m.err
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/r2c20_bench_$m.json'))
    print('A_RES=$m value', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), 'ms/step', round(d['ms_per_step'],2), 'frac', round(d['roofline']['frac'],3), d['clocks']['sm_mhz'], json.dumps(dict(list(d['roofline']['per_kernel_ms_per_step'].items())[:4])))
except Exception as e: print('ERR', e)
PY
done
