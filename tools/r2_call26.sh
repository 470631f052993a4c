#!/bin/bash
# round 2, call 26: render kernels with directed-rounding float compares instead of double (bit-exactness vs the reference kernels), kenburns_full timing
mkdir -p gpurun_out
echo "== render / kb tests"; timeout 900 python -m pytest tests/test_ref_kernels_gpu.py tests/test_kb_gpu.py tests/test_pipeline_gpu.py tests/test_bokeh_gpu.py -q -m gpu 2>&1 | tail -4 | cut -c1-300
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "== kb profile"; timeout 600 python tools/kb_profile.py 2>&1 | tail -30 | cut -c1-200
