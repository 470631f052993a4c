#!/bin/bash
# round 2, call 16: tanh-form GELU epilogue A/B (accuracy, speed, parity, bench)
mkdir -p gpurun_out
echo "== gelu_check default"; timeout 300 python tools/gelu_check.py 2>&1 | tail -5
echo "== gelu_check tanh"; CSB_GELU_FORM=tanh timeout 300 python tools/gelu_check.py 2>&1 | tail -5
echo "== parity (default)"; timeout 900 python tests/parity_full.py gpurun_out/r2c16_parity_default.json det 2>&1 | tail -12 | cut -c1-300
echo "== parity (tanh)"; CSB_GELU_FORM=tanh timeout 900 python tests/parity_full.py gpurun_out/r2c16_parity_tanh.json det 2>&1 | tail -12 | cut -c1-300
echo "== det tests under tanh"; CSB_GELU_FORM=tanh timeout 900 python -m pytest tests/test_det_gpu.py tests/test_parity_full_gpu.py tests/test_conv_gpu.py -q -m gpu -x 2>&1 | tail -5 | cut -c1-300
for form in poly tanh; do
  echo "== bench $form"
  CSB_GELU_FORM=$form timeout 600 python bench.py --steps 5 --warmup 3 --no-other --no-cpu-baseline > gpurun_out/r2c16_bench_$form.json 2> gpurun_out/r2c16_bench_$form.err; tail -2 gpurun_out/r2c16_bench_$form.err | cut -c1-300
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/r2c16_bench_$form.json'))
    print('value', d['value'], 'e2e', d['e2e']['value'], 'ms/step', d['ms_per_step'], 'frac', d['roofline']['frac'], d['clocks'])
    print('kern', json.dumps(dict(list(d['roofline']['per_kernel_ms_per_step'].items())[:6])), 'traffic', d['roofline'].get('traffic'), d['roofline'].get('traffic_over_algorithmic'))
except Exception as e: print('ERR', e)
PY
done
