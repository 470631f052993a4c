mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_isnet_gpu.py tests/test_pipeline_gpu.py tests/test_det_gpu.py -m gpu -q --timeout 600 -s > gpurun_out/pytest_new.txt 2>&1; tail -14 gpurun_out/pytest_new.txt | cut -c1-300
timeout 900 python bench.py --steps 4 --warmup 3 > gpurun_out/bench6.json 2> gpurun_out/bench6.err; tail -3 gpurun_out/bench6.err; cat gpurun_out/bench6.json | cut -c1-300
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 1500 -c 1500 --csv --log-file gpurun_out/launches_r1_full.csv python bench.py --steps 1 --warmup 3 --batch 8 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1; tail -2 gpurun_out/ncu_bench.log | cut -c1-200
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_conv_tc -s 60 -c 3 -o gpurun_out/prof_conv python tools/ncu_target.py det > gpurun_out/ncu_conv.log 2>&1; tail -2 gpurun_out/ncu_conv.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_dwconv_tile -s 20 -c 2 -o gpurun_out/prof_dw python tools/ncu_target.py det > gpurun_out/ncu_dw.log 2>&1; tail -2 gpurun_out/ncu_dw.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_splat -s 1 -c 2 -o gpurun_out/prof_splat python tools/ncu_target.py warp > gpurun_out/ncu_splat.log 2>&1; tail -2 gpurun_out/ncu_splat.log
ls -la gpurun_out/*.ncu-rep
