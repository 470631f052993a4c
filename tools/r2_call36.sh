#!/bin/bash
# round 2, call 36: k_mlp_tc with 16 epilogue warps (CSB_MLP_EG=4) vs 12
mkdir -p gpurun_out
for eg in 4 3; do
  echo "== fused mlp test EG=$eg"; CSB_MLP_EG=$eg timeout 300 python -m pytest tests/test_conv_gpu.py -q -m gpu -x -k "fused_convnext" 2>&1 | tail -3 | cut -c1-300
  echo "== mlp bench EG=$eg"; CSB_MLP_EG=$eg timeout 300 python tools/mlp_bench.py 2>&1 | tail -2
done
for eg in 4 3 4 3; do
  CSB_MLP_EG=$eg timeout 600 python bench.py --steps 10 --warmup 3 --no-other --no-cpu-baseline --stages seg > gpurun_out/r2c36_bench_$eg.json 2> gpurun_out/r2c36_bench_$eg.err; tail -2 gpurun_out/r2c36_bench_$eg.err | cut -c1-300
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/r2c36_bench_$eg.json'))
    pk=d['roofline']['per_kernel_ms_per_step']
    print('MLP_EG=$eg seg-only ms/step', round(d['ms_per_step'],2), d['clocks']['sm_mhz'], 'k_mlp_tc', pk.get('k_mlp_tc'), 'k_conv_tc', pk.get('k_conv_tc'))
except Exception as e: print('ERR', e)
PY
done
