mkdir -p gpurun_out
timeout -k 10 600 python -m pytest tests/ -x -q -m gpu 2>&1 | tail -3
timeout -k 10 900 python bench.py --steps 10 --warmup 4 --no-cpu-baseline 2>&1 | grep -v Warn | tail -1 | tee gpurun_out/bench13.json
