mkdir -p gpurun_out
timeout -k 10 600 python -m pytest tests/test_conv_gpu.py tests/test_rows_gpu.py -q -m gpu -n 4 --tb=short 2>&1 | tail -8 > gpurun_out/t_halo2.log; tail -4 gpurun_out/t_halo2.log
timeout -k 10 900 python tools/microbench.py --batch 16 --only G7,G8 > gpurun_out/microbench_halo.log 2>&1; grep "^|" gpurun_out/microbench_halo.log
timeout -k 10 900 python tools/microbench.py --batch 48 --only G7,G8 > gpurun_out/microbench_halo48.log 2>&1; grep "^|" gpurun_out/microbench_halo48.log
