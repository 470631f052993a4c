#!/bin/bash
# round 2, call 39: k_mlp_tc with three chunk buffers (C = 128: 64-unit chunks, GEMM1 two ahead) vs two
mkdir -p gpurun_out
for nb in 3 2; do
  echo "== fused mlp test NB=$nb"; CSB_MLP_NB=$nb timeout 300 python -m pytest tests/test_conv_gpu.py -q -m gpu -x -k "fused_convnext" 2>&1 | tail -4 | cut -c1-300
  echo "== mlp bench NB=$nb"; CSB_MLP_NB=$nb timeout 300 python tools/mlp_bench.py 2>&1 | tail -2
done
echo "== NB=3 pair"; CSB_MLP_PAIR=1 timeout 300 python -m pytest tests/test_conv_gpu.py -q -m gpu -x -k "fused_convnext" 2>&1 | tail -2 | cut -c1-300; CSB_MLP_PAIR=1 timeout 300 python tools/mlp_bench.py 2>&1 | tail -2
echo "== det tests"; timeout 600 python -m pytest tests/test_det_gpu.py tests/test_parity_full_gpu.py -q -m gpu -x 2>&1 | tail -3 | cut -c1-300
for nb in 3 2 3 2; do
  CSB_MLP_NB=$nb timeout 600 python bench.py --steps 10 --warmup 3 --no-other --no-cpu-baseline > gpurun_out/r2c39_bench_$nb.json 2> gpurun_out/r2c39_bench_$nb.err; tail -2 gpurun_out/r2c39_bench_$nb.err | cut -c1-300
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/r2c39_bench_$nb.json'))
    pk=d['roofline']['per_kernel_ms_per_step']
    print('MLP_NB=$nb value', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), 'ms/step', round(d['ms_per_step'],2), d['clocks']['sm_mhz'], 'k_mlp_tc', pk.get('k_mlp_tc'), 'k_conv_tc', pk.get('k_conv_tc'))
except Exception as e: print('ERR', e)
PY
done
