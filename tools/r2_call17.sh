#!/bin/bash
# round 2, call 17: fp32-output epilogue (coalesced), Lanczos tail, GELU form A/B per layer
mkdir -p gpurun_out
echo "== tests"; timeout 900 python -m pytest tests/test_conv_gpu.py tests/test_leres_gpu.py tests/test_det_gpu.py tests/test_halo_gpu.py -q -m gpu -x 2>&1 | tail -6 | cut -c1-300
for form in poly tanh; do
  echo "== layer profile $form"
  CSB_GELU_FORM=$form timeout 600 python tools/layer_profile.py 32 gpurun_out/r2c17_layers_$form.json > gpurun_out/r2c17_layers_$form.md 2>&1; grep -E "==|act3|->169" gpurun_out/r2c17_layers_$form.md | cut -c1-160
done
echo "== bench x2 alternating"
for form in tanh poly tanh poly; do
  CSB_GELU_FORM=$form timeout 600 python bench.py --steps 10 --warmup 3 --no-other --no-cpu-baseline > gpurun_out/r2c17_bench_$form.json 2> gpurun_out/r2c17_bench_$form.err
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/r2c17_bench_$form.json'))
    print('$form value', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), 'ms/step', round(d['ms_per_step'],2), 'frac', round(d['roofline']['frac'],3), d['clocks']['sm_mhz'], json.dumps(dict(list(d['roofline']['per_kernel_ms_per_step'].items())[:4])))
except Exception as e: print('ERR', e)
PY
done
