#!/bin/bash
# round 2, call 32: LeReS stem on the halo kernel (input padded to 64 channels) vs the per-tap kernel (16); fused-MLP bf16 test
mkdir -p gpurun_out
echo "== tests"; timeout 900 python -m pytest tests/test_leres_gpu.py tests/test_parity_full_gpu.py tests/test_conv_gpu.py -q -m gpu -x -k "not a_resident" 2>&1 | tail -3 | cut -c1-300
for cp in 64 16 64 16; do
  CSB_LERES_STEM_CP=$cp timeout 600 python bench.py --steps 10 --warmup 3 --no-other --no-cpu-baseline --stages depth > gpurun_out/r2c32_bench_$cp.json 2> gpurun_out/r2c32_bench_$cp.err; tail -2 gpurun_out/r2c32_bench_$cp.err | cut -c1-300
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/r2c32_bench_$cp.json'))
    pk=d['roofline']['per_kernel_ms_per_step']
    print('STEM_CP=$cp depth-only ms/step', round(d['ms_per_step'],2), d['clocks']['sm_mhz'], json.dumps({k:pk[k] for k in list(pk)[:6]}), 'prep', pk.get('k_image_prep_s2d'))
except Exception as e: print('ERR', e)
PY
done
CSB_LERES_STEM_CP=16 timeout 300 python tests/parity_full.py gpurun_out/r2c32_parity16.json 2>&1 | grep -A8 '"leres_640"' | head -12
timeout 300 python tests/parity_full.py gpurun_out/r2c32_parity64.json 2>&1 | grep -A8 '"leres_640"' | head -12
