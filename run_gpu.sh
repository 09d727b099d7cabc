mkdir -p gpurun_out
timeout -k 10 1500 python -m pytest tests/test_modules_gpu.py tests/test_rows_gpu.py tests/test_step_gpu.py tests/test_graph_gpu.py -q -m gpu -n 4 --tb=short 2>&1 | grep -v "^frame\|python()\|Warning\|warnings.html\|detach()" | tail -40 > gpurun_out/t_net.log; tail -12 gpurun_out/t_net.log
timeout -k 10 900 python bench.py --steps 10 --warmup 4 --no-cpu-baseline 2>&1 | grep -v Warn | tail -1 | tee gpurun_out/bench7.json
