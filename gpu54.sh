#!/bin/bash
mkdir -p gpurun_out
echo "== tc"; timeout 400 python tools/zoe_bench.py 16 gpurun_out/zoe54_tc.json 2>&1 | grep -v Warn | tail -13
echo "== mma.sync"; CSB_ATTN_TC=0 timeout 400 python tools/zoe_bench.py 16 gpurun_out/zoe54_mma.json 2>&1 | grep -v Warn | tail -13
