mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; tail -5 gpurun_out/bench_n2.err | cut -c1-300; cat gpurun_out/bench_n2.json | cut -c1-400
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 1 --warmup 0 --stages warp > gpurun_out/bench_n2_ref.json 2> gpurun_out/bench_n2_ref.err; tail -2 gpurun_out/bench_n2_ref.err | cut -c1-200; cat gpurun_out/bench_n2_ref.json | cut -c1-200
