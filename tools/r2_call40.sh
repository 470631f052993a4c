#!/bin/bash
# round 2, call 40: does the mbarrier suspend hint lengthen the hand-offs of k_mlp_tc?  (hint 4000 ns = default build, 200 ns, none)
echo "== hint 4000"; timeout 300 python tools/mlp_bench.py 2>&1 | tail -2
echo "== hint 200"; CSB200_LIB=$PWD/cartoonsegmentation_b200/libcsb200_h200.so timeout 300 python tools/mlp_bench.py 2>&1 | tail -2
echo "== no hint"; CSB200_LIB=$PWD/cartoonsegmentation_b200/libcsb200_nohint.so timeout 300 python tools/mlp_bench.py 2>&1 | tail -2
echo "== gelu_check hint 4000"; timeout 300 python tools/gelu_check.py 2>&1 | tail -3
echo "== gelu_check hint 200"; CSB200_LIB=$PWD/cartoonsegmentation_b200/libcsb200_h200.so timeout 300 python tools/gelu_check.py 2>&1 | tail -3
