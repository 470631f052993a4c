mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_conv_gpu.py tests/test_leres_gpu.py tests/test_pipeline_gpu.py -m gpu -q --timeout 600 -s > gpurun_out/pytest_leres.txt 2>&1; tail -25 gpurun_out/pytest_leres.txt | cut -c1-300
timeout 600 python tools/det_profile.py 32 gpurun_out/det_profile.json 2>&1 | tail -25
timeout 600 python tools/conv_bench.py gpurun_out/conv_bench3.json 2>&1 | grep -E "fc1|stem" 
