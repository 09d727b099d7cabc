mkdir -p gpurun_out
timeout -k 10 900 python -m pytest tests/test_modules_gpu.py -q -m gpu -k "text_encoder" --tb=short -rP 2>&1 | grep -v "Warning\|warnings.html" | tail -30 > gpurun_out/t_txt.log; tail -30 gpurun_out/t_txt.log
timeout -k 10 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | grep -v Warn | tail -1 | tee gpurun_out/bench3.json
