#!/bin/bash
# round 2, call 21: mbarrier try_wait suspend-time hint A/B (libcsb200.so = 4 us hint, libcsb200_nohint.so = round-1 polling loop)
mkdir -p gpurun_out
NOHINT=$PWD/cartoonsegmentation_b200/libcsb200_nohint.so
echo "== tests (hint build)"; timeout 1200 python -m pytest tests/test_conv_gpu.py tests/test_halo_gpu.py tests/test_dw_halo_gpu.py tests/test_zoe_dpt_gpu.py tests/test_det_gpu.py -q -m gpu -x 2>&1 | tail -4 | cut -c1-300
echo "== gelu_check hint"; timeout 300 python tools/gelu_check.py 2>&1 | tail -3
echo "== gelu_check nohint"; CSB200_LIB=$NOHINT timeout 300 python tools/gelu_check.py 2>&1 | tail -3
for v in hint nohint hint nohint; do
  if [ $v = nohint ]; then export CSB200_LIB=$NOHINT; else unset CSB200_LIB; fi
  timeout 600 python bench.py --steps 10 --warmup 3 --no-other --no-cpu-baseline > gpurun_out/r2c21_bench_$v.json 2> gpurun_out/r2c21_bench_$v.err
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/r2c21_bench_$v.json'))
    print('$v value', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), 'ms/step', round(d['ms_per_step'],2), 'frac', round(d['roofline']['frac'],3), d['clocks']['sm_mhz'], json.dumps(dict(list(d['roofline']['per_kernel_ms_per_step'].items())[:4])))
except Exception as e: print('ERR', e)
PY
done
unset CSB200_LIB
echo "== zoe bench hint vs nohint"
for v in hint nohint; do
  if [ $v = nohint ]; then export CSB200_LIB=$NOHINT; else unset CSB200_LIB; fi
  timeout 600 python bench.py --depth zoe --steps 3 --warmup 3 --no-other --no-cpu-baseline > gpurun_out/r2c21_zoe_$v.json 2> gpurun_out/r2c21_zoe_$v.err
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/r2c21_zoe_$v.json'))
    print('zoe $v value', round(d['value'],1), 'ms/step', round(d['ms_per_step'],2), d['clocks']['sm_mhz'], json.dumps(dict(list(d['roofline']['per_kernel_ms_per_step'].items())[:4])))
except Exception as e: print('ERR', e)
PY
done
