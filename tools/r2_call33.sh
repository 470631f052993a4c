#!/bin/bash
# round 2, call 33: specialised P = 2 prep kernel + LeReS stem on the halo kernel (CP 64) vs per-tap (CP 16)
mkdir -p gpurun_out
echo "== tests"; timeout 900 python -m pytest tests/test_leres_gpu.py tests/test_conv_gpu.py -q -m gpu -x 2>&1 | tail -3 | cut -c1-300
for cp in 64 16 64 16; do
  CSB_LERES_STEM_CP=$cp timeout 600 python bench.py --steps 10 --warmup 3 --no-other --no-cpu-baseline --stages depth > gpurun_out/r2c33_bench_$cp.json 2> gpurun_out/r2c33_bench_$cp.err; tail -2 gpurun_out/r2c33_bench_$cp.err | cut -c1-300
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/r2c33_bench_$cp.json'))
    pk=d['roofline']['per_kernel_ms_per_step']
    print('STEM_CP=$cp depth-only ms/step', round(d['ms_per_step'],2), d['clocks']['sm_mhz'], 'conv_tc', pk.get('k_conv_tc'), 'halo', pk.get('k_conv_halo'), 'prep', pk.get('k_image_prep'))
except Exception as e: print('ERR', e)
PY
done
