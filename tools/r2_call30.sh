#!/bin/bash
# round 2, call 30: compute-sanitizer memcheck over the kernels added / changed this round (small shapes)
mkdir -p gpurun_out
run() { echo "== memcheck: $*"; timeout 900 compute-sanitizer --tool memcheck --error-exitcode 99 --print-limit 5 "$@" > gpurun_out/r2c30_san.log 2>&1; rc=$?; echo "rc=$rc"; grep -E "ERROR SUMMARY|passed|failed|Invalid|out of bounds|misaligned" gpurun_out/r2c30_san.log | tail -4 | cut -c1-200; cat gpurun_out/r2c30_san.log >> gpurun_out/r2c30_sanitizer_all.log; }
run python -m pytest tests/test_conv_gpu.py -q -m gpu -x -k "fused_convnext and (33 or 24)"
run python -m pytest tests/test_conv_gpu.py -q -m gpu -x -k "cta_pair and (24-40 or 33-47)"
run python -m pytest tests/test_leres_gpu.py -q -m gpu -x -k "lanczos and (96 or 32-32)"
run python -m pytest tests/test_pipeline_gpu.py -q -m gpu -x -k "net_io"
run python -m pytest tests/test_det_gpu.py -q -m gpu -x -k "chunked_topk and 1280"
run python -m pytest tests/test_halo_gpu.py -q -m gpu -x -k "not slow" --maxfail=1 -k "64"
run python -c "import __graft_entry__ as g; g.smoke()"
grep -c "ERROR SUMMARY: 0 errors" gpurun_out/r2c30_sanitizer_all.log
