mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_conv_gpu.py tests/test_det_gpu.py -m gpu -q --timeout 600 -s > gpurun_out/pytest_det.txt 2>&1; tail -12 gpurun_out/pytest_det.txt | cut -c1-300
timeout 600 python tools/conv_bench.py gpurun_out/conv_bench2.json 2>&1 | tail -20
