#!/bin/bash
# round 2, call 46: residual L2 prefetch in k_mlp_tc on / off (twice each, alternating)
for v in 1 0 1 0; do echo "== RES_PF=$v"; CSB_MLP_RES_PF=$v timeout 300 python tools/mlp_bench.py 2>&1 | tail -2; done
