mkdir -p gpurun_out
timeout -k 10 600 python -m pytest tests/test_post_gpu.py tests/test_modules_gpu.py tests/test_step_gpu.py tests/test_graph_gpu.py -q -m gpu --tb=short 2>&1 | grep -v "Warning\|warnings.html\|detach()" | tail -30
