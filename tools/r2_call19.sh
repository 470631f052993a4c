#!/bin/bash
# round 2, call 19: ncu --set full of the stage-3 fc1 GEMM (CTA-pair kernel, tanh GELU / no activation) and of the halo depthwise kernel, reduced to CSV
# on the box (the reports exceed the 64 MiB return limit); full bench
mkdir -p gpurun_out /tmp/ncu
cap() {  # name kernel-regex skip cmd...
  local name=$1 rx=$2 skip=$3; shift 3
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:$rx -s $skip -c 1 -f -o /tmp/ncu/$name "$@" 2>&1 | tail -1
  ncu -i /tmp/ncu/$name.ncu-rep --page raw --csv > gpurun_out/r2c19_${name}_raw.csv 2>/dev/null
  ncu -i /tmp/ncu/$name.ncu-rep --page source --csv 2>/dev/null | gzip > gpurun_out/r2c19_${name}_source.csv.gz
  ncu -i /tmp/ncu/$name.ncu-rep --page details 2>/dev/null | grep -E "Duration|Throughput|Pipe|Issue|Eligible|Stall|stall|Registers|Theoretical Occ|Achieved Occ|L2 Cache|DRAM|Executed Ipc|No Eligible|One or More" | head -60 > gpurun_out/r2c19_${name}_details.txt
  ls -la gpurun_out/r2c19_${name}_*
}
cap fc1 k_conv_tc 2 python tools/conv_one.py 32 64 64 512 2048 1 gelu
cap fc1_noact k_conv_tc 2 python tools/conv_one.py 32 64 64 512 2048 1 none
cap dw k_dwconv_halo 1 python tools/dw_one.py
echo "== bench"
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2c19_bench.json 2> gpurun_out/r2c19_bench.err; tail -3 gpurun_out/r2c19_bench.err | cut -c1-400
python - <<'PY'
import json
try:
    d=json.load(open('gpurun_out/r2c19_bench.json'))
    print('value', d['value'], 'e2e', d['e2e']['value'], 'ms/step', d['ms_per_step'], 'frac', d['roofline']['frac'], 'launches', d['gpu_launches'], d['clocks'])
    print('cpu', d.get('cpu_baseline',{}).get('value'))
    for k,v in (d.get('other_workloads') or {}).items(): print(k, json.dumps({a:b for a,b in v.items() if a not in('api','per_kernel_ms_profiled')})[:400])
except Exception as e: print('ERR', e)
PY
du -sh gpurun_out
