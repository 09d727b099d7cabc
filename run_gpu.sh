mkdir -p gpurun_out
timeout -k 10 900 python -m pytest tests/test_graph_gpu.py -q -m gpu --tb=short -x 2>&1 | grep -v "Warning\|warnings.html" | tail -40 > gpurun_out/t_graph.log; tail -40 gpurun_out/t_graph.log
timeout -k 10 900 python bench.py --steps 10 --warmup 4 --no-cpu-baseline 2>&1 | grep -v Warn | tail -3 | tee gpurun_out/bench4.json
