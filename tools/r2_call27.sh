#!/bin/bash
# round 2, call 27: fused ConvNeXt MLP kernel (k_mlp_tc): bit-identity test, detector tests, step A/B
mkdir -p gpurun_out
echo "== fused mlp test"; timeout 300 python -m pytest tests/test_conv_gpu.py -q -m gpu -x -k "fused_convnext" 2>&1 | tail -12 | cut -c1-400
echo "== det tests"; timeout 600 python -m pytest tests/test_det_gpu.py tests/test_parity_full_gpu.py -q -m gpu -x 2>&1 | tail -4 | cut -c1-300
for m in 1 0 1 0; do
  CSB_FUSE_MLP=$m timeout 600 python bench.py --steps 10 --warmup 3 --no-other --no-cpu-baseline > gpurun_out/r2c27_bench_$m.json 2> gpurun_out/r2c27_bench_$m.err; tail -2 gpurun_out/r2c27_bench_$m.err | cut -c1-300
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/r2c27_bench_$m.json'))
    print('FUSE_MLP=$m value', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), 'ms/step', round(d['ms_per_step'],2), 'frac', round(d['roofline']['frac'],3), d['clocks']['sm_mhz'], json.dumps(dict(list(d['roofline']['per_kernel_ms_per_step'].items())[:5])))
except Exception as e: print('ERR', e)
PY
done
