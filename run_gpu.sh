mkdir -p gpurun_out
timeout -k 10 600 python -m pytest tests/ -x -q -m gpu --tb=short 2>&1 | grep -v "Warning\|warnings.html\|detach()" | tail -4
