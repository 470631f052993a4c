#!/bin/bash
# round 2, call 43: k_adjp_cover16 at 126 registers (2 CTAs / SM) with 16 loads in flight per thread (was 228 registers, 1 CTA / SM, 8 loads)
mkdir -p gpurun_out
echo "== tests"; timeout 600 python -m pytest tests/test_pipeline_gpu.py tests/test_kb_gpu.py -q -m gpu 2>&1 | tail -3 | cut -c1-300
for i in 1 2; do
  timeout 600 python bench.py --steps 10 --warmup 3 --no-other --no-cpu-baseline > gpurun_out/r2c43_bench.json 2> gpurun_out/r2c43_bench.err; tail -2 gpurun_out/r2c43_bench.err | cut -c1-300
  python - <<'PY'
import json
d=json.load(open('gpurun_out/r2c43_bench.json'))
pk=d['roofline']['per_kernel_ms_per_step']
print('value', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), 'ms/step', round(d['ms_per_step'],2), d['clocks']['sm_mhz'], 'k_adjp_cover', pk.get('k_adjp_cover'), 'k_adjp_rel', pk.get('k_adjp_rel'), 'k_mask_tail', pk.get('k_mask_tail'))
PY
done
