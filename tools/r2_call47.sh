#!/bin/bash
# round 2, call 47: dispatch check after trimming the experimental k_mlp_tc instantiations to fp16
timeout 200 python -m pytest tests/test_conv_gpu.py -q -m gpu -x -k "fused_convnext" 2>&1 | tail -2 | cut -c1-200
CSB_MLP_EG=4 timeout 200 python -m pytest tests/test_conv_gpu.py -q -m gpu -x -k "fused_convnext" 2>&1 | tail -2 | cut -c1-200
CSB_MLP_PAIR=1 timeout 200 python -m pytest tests/test_conv_gpu.py -q -m gpu -x -k "fused_convnext" 2>&1 | tail -2 | cut -c1-200
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
