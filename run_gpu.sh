mkdir -p gpurun_out
timeout -k 10 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus 2 --steps 10 --warmup 4 --no-cpu-baseline > gpurun_out/bench_2gpu.log 2>&1
grep '"metric"' gpurun_out/bench_2gpu.log | head -1 | cut -c1-330; tail -2 gpurun_out/bench_2gpu.log | cut -c1-200
