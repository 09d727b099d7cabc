mkdir -p gpurun_out
timeout -k 10 1500 python -m pytest tests/test_modules_gpu.py -q -m gpu -n 4 --tb=line 2>&1 | tail -30 > gpurun_out/test_modules_gpu.log
echo "== modules"; tail -30 gpurun_out/test_modules_gpu.log
timeout -k 10 1500 python -m pytest tests/test_step_gpu.py -q -m gpu -n 2 --tb=short -rP 2>&1 | grep -v "^frame\|python()\|Warning" | tail -80 > gpurun_out/test_step_gpu.log
echo "== step"; tail -80 gpurun_out/test_step_gpu.log
