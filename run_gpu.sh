mkdir -p gpurun_out
(time timeout -k 10 600 python -m pytest tests -q -m gpu -x --tb=short 2>&1 | grep -v "Warning\|warnings.html\|detach()" | tail -12) > gpurun_out/t_all.log 2>&1
tail -6 gpurun_out/t_all.log
timeout -k 10 300 python bench.py --steps 10 --warmup 4 --no-cpu-baseline 2>gpurun_out/bench38.err | tail -1 > gpurun_out/bench38.json
DWC_STY_STREAM=0 timeout -k 10 300 python bench.py --steps 10 --warmup 4 --no-cpu-baseline 2>gpurun_out/bench39.err | tail -1 > gpurun_out/bench39.json
cut -c1-200 gpurun_out/bench38.json gpurun_out/bench39.json; tail -2 gpurun_out/bench38.err
