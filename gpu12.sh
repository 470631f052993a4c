mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 -s > gpurun_out/pytest_all.txt 2>&1; grep -E "passed|failed|ZoeDepth head|error" gpurun_out/pytest_all.txt | tail -8 | cut -c1-300
timeout 600 python tools/det_profile.py 32 gpurun_out/det_profile5.json 2>&1 | tail -11
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench8.json 2> gpurun_out/bench8.err; tail -3 gpurun_out/bench8.err; cat gpurun_out/bench8.json | cut -c1-300
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
