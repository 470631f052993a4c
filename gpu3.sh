mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kb_gpu.py tests/test_ref_kernels_gpu.py -m gpu -q --timeout 600 > gpurun_out/pytest_kb.txt 2>&1; tail -4 gpurun_out/pytest_kb.txt
timeout 1200 python -m pytest tests/test_det_gpu.py -m gpu -q --timeout 600 -s > gpurun_out/pytest_det.txt 2>&1; tail -40 gpurun_out/pytest_det.txt
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench2.json 2> gpurun_out/bench2.err; tail -3 gpurun_out/bench2.err; cat gpurun_out/bench2.json
