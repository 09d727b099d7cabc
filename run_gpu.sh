mkdir -p gpurun_out
timeout -k 5 170 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 4 > gpurun_out/bench_2gpu.log 2>&1; echo "rc=$?"; grep -v "Warn\|warn" gpurun_out/bench_2gpu.log | tail -3
