mkdir -p gpurun_out
timeout -k 10 900 python -m pytest tests/test_conv_gpu.py tests/test_rows_gpu.py tests/test_post_gpu.py -q -m gpu -n 4 --tb=short 2>&1 | tail -15 > gpurun_out/t_units.log; tail -5 gpurun_out/t_units.log
timeout -k 10 1500 python -m pytest tests/test_modules_gpu.py tests/test_step_gpu.py tests/test_graph_gpu.py -q -m gpu -n 4 --tb=short 2>&1 | grep -v "^frame\|python()\|Warning\|warnings.html\|detach()" | tail -30 > gpurun_out/t_net.log; tail -8 gpurun_out/t_net.log
timeout -k 10 900 python tools/microbench.py --batch 16 --out gpurun_out/microbench_b16.md > gpurun_out/microbench.log 2>&1; tail -5 gpurun_out/microbench.log
timeout -k 10 900 python bench.py --steps 10 --warmup 4 --no-cpu-baseline 2>&1 | grep -v Warn | tail -1 | tee gpurun_out/bench5.json
