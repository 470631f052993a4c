mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kb_gpu.py tests/test_ref_kernels_gpu.py tests/test_conv_gpu.py -m gpu -q --timeout 600 > gpurun_out/pytest_kb.txt 2>&1; tail -4 gpurun_out/pytest_kb.txt
timeout 1200 python -m pytest tests/test_det_gpu.py tests/test_pipeline_gpu.py -m gpu -q --timeout 600 -s > gpurun_out/pytest_det.txt 2>&1; tail -40 gpurun_out/pytest_det.txt | cut -c1-300
timeout 600 python tools/conv_bench.py gpurun_out/conv_bench.json 2>&1 | tail -20
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench3.json 2> gpurun_out/bench3.err; tail -3 gpurun_out/bench3.err; cat gpurun_out/bench3.json | cut -c1-400
