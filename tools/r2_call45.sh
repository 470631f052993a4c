#!/bin/bash
# round 2, call 45: k_mlp_tc with colsum / b1 staged in shared memory + residual L2 prefetch
mkdir -p gpurun_out
for v in 1 0; do
  echo "== fused mlp test VEC_SMEM=$v"; CSB_MLP_VEC_SMEM=$v timeout 300 python -m pytest tests/test_conv_gpu.py -q -m gpu -x -k "fused_convnext" 2>&1 | tail -3 | cut -c1-300
  echo "== mlp bench VEC_SMEM=$v"; CSB_MLP_VEC_SMEM=$v timeout 300 python tools/mlp_bench.py 2>&1 | tail -2
done
echo "== det tests"; timeout 600 python -m pytest tests/test_det_gpu.py tests/test_parity_full_gpu.py -q -m gpu -x 2>&1 | tail -3 | cut -c1-300
