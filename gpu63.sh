#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_attention_tc -s 2 -c 1 -o gpurun_out/prof_attn python tools/attn_one.py 8 1765 > gpurun_out/ncu63.log 2>&1; tail -2 gpurun_out/ncu63.log
ls -la gpurun_out/prof_attn.ncu-rep
