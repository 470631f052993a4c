#!/bin/bash
# round 2, call 18: ncu --set full of the stage-3 fc1 GEMM (CTA-pair kernel, tanh GELU) and of the halo depthwise kernel; full GPU tests; full bench
mkdir -p gpurun_out
echo "== ncu fc1"; timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_conv_tc -s 2 -c 1 -f -o gpurun_out/r2c18_fc1 python tools/conv_one.py 32 64 64 512 2048 1 gelu 2>&1 | tail -2
echo "== ncu fc1 no act"; timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_conv_tc -s 2 -c 1 -f -o gpurun_out/r2c18_fc1_noact python tools/conv_one.py 32 64 64 512 2048 1 none 2>&1 | tail -2
echo "== ncu dw"; timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_dwconv_halo -s 1 -c 1 -f -o gpurun_out/r2c18_dw python tools/dw_one.py 2>&1 | tail -2
ls -la gpurun_out/*.ncu-rep
echo "== all gpu tests"; timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r2c18_tests.log 2>&1; tail -4 gpurun_out/r2c18_tests.log | cut -c1-300
echo "== bench"
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2c18_bench.json 2> gpurun_out/r2c18_bench.err; tail -3 gpurun_out/r2c18_bench.err | cut -c1-400
python - <<'PY'
import json
try:
    d=json.load(open('gpurun_out/r2c18_bench.json'))
    print('value', d['value'], 'e2e', d['e2e']['value'], 'ms/step', d['ms_per_step'], 'frac', d['roofline']['frac'], 'launches', d['gpu_launches'], d['clocks'])
    print('cpu', d.get('cpu_baseline',{}).get('value'))
    for k,v in (d.get('other_workloads') or {}).items(): print(k, json.dumps({a:b for a,b in v.items() if a not in('api','per_kernel_ms_profiled')})[:500])
except Exception as e: print('ERR', e)
PY
