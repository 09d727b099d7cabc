mkdir -p gpurun_out
timeout -k 10 900 python -m pytest tests/test_configs_gpu.py -q -m gpu -n 3 --tb=short -rP 2>&1 | grep -v "^frame\|python()\|Warning\|warnings.html\|detach()" | tail -30 > gpurun_out/t_cfg.log; tail -30 gpurun_out/t_cfg.log
