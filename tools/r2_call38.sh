#!/bin/bash
# round 2, call 38: k_mlp_tc on CTA pairs (cta_group::2, half a weight chunk per CTA) vs single CTAs
mkdir -p gpurun_out
for pr in 1 0; do
  echo "== fused mlp test PAIR=$pr"; CSB_MLP_PAIR=$pr timeout 300 python -m pytest tests/test_conv_gpu.py -q -m gpu -x -k "fused_convnext" 2>&1 | tail -4 | cut -c1-300
  echo "== mlp bench PAIR=$pr"; CSB_MLP_PAIR=$pr timeout 300 python tools/mlp_bench.py 2>&1 | tail -2
done
echo "== det tests (pair)"; timeout 600 python -m pytest tests/test_det_gpu.py tests/test_parity_full_gpu.py -q -m gpu -x 2>&1 | tail -3 | cut -c1-300
for pr in 1 0 1 0; do
  CSB_MLP_PAIR=$pr timeout 600 python bench.py --steps 10 --warmup 3 --no-other --no-cpu-baseline > gpurun_out/r2c38_bench_$pr.json 2> gpurun_out/r2c38_bench_$pr.err; tail -2 gpurun_out/r2c38_bench_$pr.err | cut -c1-300
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/r2c38_bench_$pr.json'))
    pk=d['roofline']['per_kernel_ms_per_step']
    print('MLP_PAIR=$pr value', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), 'ms/step', round(d['ms_per_step'],2), d['clocks']['sm_mhz'], 'k_mlp_tc', pk.get('k_mlp_tc'), 'k_conv_tc', pk.get('k_conv_tc'))
except Exception as e: print('ERR', e)
PY
done
