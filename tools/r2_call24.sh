#!/bin/bash
# round 2, call 24 (2 GPUs): full GPU suite, smoke, reference arm, 2-rank bench
mkdir -p gpurun_out
echo "== all gpu tests"; ( time timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r2c24_tests.log 2>&1 ) 2>&1 | grep real; tail -3 gpurun_out/r2c24_tests.log | cut -c1-300
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "== reference arm"; timeout 600 python bench.py --impl reference --steps 1 --warmup 1 2>/dev/null | cut -c1-600
echo "== 2-rank bench"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r2c24_bench2.json 2> gpurun_out/r2c24_bench2.err; tail -2 gpurun_out/r2c24_bench2.err | cut -c1-300
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/r2c24_bench2.json').read().strip().splitlines()[-1])
    print('N=2 value', d['value'], 'e2e', d['e2e']['value'], 'ms/step', d['ms_per_step'], 'n_gpus', d['n_gpus'], d['clocks'])
except Exception as e: print('ERR', e)
PY
