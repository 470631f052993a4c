#!/bin/bash
mkdir -p gpurun_out
for sub in 32 16; do
CSB_SEG_SUB=$sub timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/bench12_$sub.json 2> gpurun_out/bench12_$sub.err; tail -2 gpurun_out/bench12_$sub.err; python - <<PY
import json
d=json.load(open('gpurun_out/bench12_$sub.json'))
print($sub, d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'], d['gpu_launches'])
PY
done
