#!/bin/bash
# round 2, call 28: fused MLP, 16- vs 32-column EPI1 tasks
mkdir -p gpurun_out
for s in 1 0; do
  echo "== fused mlp test SUB16=$s"; CSB_MLP_SUB16=$s timeout 300 python -m pytest tests/test_conv_gpu.py -q -m gpu -x -k "fused_convnext" 2>&1 | tail -3 | cut -c1-300
  echo "== mlp bench SUB16=$s"; CSB_MLP_SUB16=$s timeout 300 python tools/mlp_bench.py 2>&1 | tail -3
done
echo "== poly form"; CSB_GELU_FORM=poly timeout 300 python -m pytest tests/test_conv_gpu.py -q -m gpu -x -k "fused_convnext" 2>&1 | tail -3 | cut -c1-300
for m in 1 0 1 0; do
  CSB_FUSE_MLP=$m timeout 600 python bench.py --steps 10 --warmup 3 --no-other --no-cpu-baseline > gpurun_out/r2c28_bench_$m.json 2> gpurun_out/r2c28_bench_$m.err; tail -2 gpurun_out/r2c28_bench_$m.err | cut -c1-300
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/r2c28_bench_$m.json'))
    print('FUSE_MLP=$m value', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), 'ms/step', round(d['ms_per_step'],2), 'frac', round(d['roofline']['frac'],3), d['clocks']['sm_mhz'], json.dumps(dict(list(d['roofline']['per_kernel_ms_per_step'].items())[:5])))
except Exception as e: print('ERR', e)
PY
done
