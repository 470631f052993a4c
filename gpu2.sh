mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 600 -k "not conv" > gpurun_out/pytest_kb.txt 2>&1; tail -5 gpurun_out/pytest_kb.txt
timeout 900 python -m pytest tests/test_conv_gpu.py -m gpu -q --timeout 120 -x > gpurun_out/pytest_conv.txt 2>&1; tail -30 gpurun_out/pytest_conv.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
