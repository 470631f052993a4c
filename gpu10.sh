mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/pytest_all.txt 2>&1; tail -6 gpurun_out/pytest_all.txt | cut -c1-300
timeout 600 python tools/conv_bench.py gpurun_out/conv_bench4.json 2>&1 | tail -17
timeout 600 python tools/det_profile.py 32 gpurun_out/det_profile4.json 2>&1 | tail -12
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench7.json 2> gpurun_out/bench7.err; tail -3 gpurun_out/bench7.err; cat gpurun_out/bench7.json | cut -c1-300
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench7_ref.json 2> gpurun_out/bench7_ref.err; tail -2 gpurun_out/bench7_ref.err; cat gpurun_out/bench7_ref.json | cut -c1-250
