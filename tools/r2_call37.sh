#!/bin/bash
# round 2, call 37: what bounds k_mlp_tc -- the kernel with parts of EPI1 removed (CSB_MLP_DIAG, wrong results, timing only)
for d in 0 1 2 3 7; do echo "== DIAG=$d (1 no GELU, 2 no LN fold, 4 no smem store)"; CSB_MLP_DIAG=$d timeout 300 python tools/mlp_bench.py 2>&1 | tail -2; done
