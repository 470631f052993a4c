#!/bin/bash
# round 2, call 44: ncu --set full of k_mlp_tc (C = 128 and C = 256 stage shapes), reduced to CSV on the box
mkdir -p gpurun_out /tmp/ncu
cap() {
  local name=$1; shift
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_mlp_tc -s 2 -c 1 -f -o /tmp/ncu/$name "$@" 2>&1 | tail -1
  ncu -i /tmp/ncu/$name.ncu-rep --page raw --csv > gpurun_out/r2c44_${name}_raw.csv 2>/dev/null
  ncu -i /tmp/ncu/$name.ncu-rep --page source --csv 2>/dev/null | gzip > gpurun_out/r2c44_${name}_source.csv.gz
  ncu -i /tmp/ncu/$name.ncu-rep --page details 2>/dev/null | grep -E "Duration|Throughput|Pipe|Issue|Eligible|Registers|Theoretical Occ|DRAM|Executed Ipc|No Eligible|One or More" | head -40 > gpurun_out/r2c44_${name}_details.txt
  ls -la gpurun_out/r2c44_${name}_*
}
cap mlp128 python tools/mlp_one.py 128
cap mlp256 python tools/mlp_one.py 256
