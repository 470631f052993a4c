mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/pytest_all.txt 2>&1; tail -12 gpurun_out/pytest_all.txt | cut -c1-300
timeout 600 python tools/det_profile.py 32 gpurun_out/det_profile2.json 2>&1 | tail -14
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/bench4.json 2> gpurun_out/bench4.err; tail -5 gpurun_out/bench4.err; cat gpurun_out/bench4.json | cut -c1-3000
