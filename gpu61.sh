#!/bin/bash
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -x -q -m gpu --durations=8 > gpurun_out/t61.log 2>&1 ) 2>&1 | grep real; tail -16 gpurun_out/t61.log | cut -c1-200
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 2600 -c 2400 --csv --log-file gpurun_out/launches_r1s2_full.csv python bench.py --steps 1 --warmup 3 --batch 8 --no-cpu-baseline --no-other > gpurun_out/ncu61.log 2>&1; tail -2 gpurun_out/ncu61.log | cut -c1-200; wc -l gpurun_out/launches_r1s2_full.csv
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
