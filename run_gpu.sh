mkdir -p gpurun_out
timeout -k 10 1200 python -m pytest tests/test_curve_gpu.py -q -m gpu --tb=short -rP 2>&1 | grep -v "^frame\|python()\|Warning\|warnings.html\|detach()" | tail -14 > gpurun_out/t_curve.log; tail -14 gpurun_out/t_curve.log
timeout -k 10 900 python -m pytest tests/test_post_gpu.py tests/test_step_gpu.py tests/test_graph_gpu.py tests/test_modules_gpu.py -q -m gpu -n 4 --tb=short 2>&1 | grep -v "^frame\|python()\|Warning\|warnings.html\|detach()" | tail -8 > gpurun_out/t_net.log; tail -5 gpurun_out/t_net.log
timeout -k 10 900 python bench.py --steps 10 --warmup 4 --no-cpu-baseline 2>&1 | grep -v Warn | tail -1 | tee gpurun_out/bench11.json
