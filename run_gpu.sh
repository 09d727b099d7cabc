mkdir -p gpurun_out
timeout -k 10 900 python -m pytest tests/test_post_gpu.py tests/test_modules_gpu.py tests/test_step_gpu.py tests/test_graph_gpu.py -q -m gpu -n 4 --tb=short 2>&1 | grep -v "^frame\|python()\|Warning\|warnings.html\|detach()" | tail -12 > gpurun_out/t_net.log; tail -5 gpurun_out/t_net.log
timeout -k 10 900 python tools/microbench.py --batch 48 --only NONE > gpurun_out/microbench_norm.log 2>&1; grep "^|" gpurun_out/microbench_norm.log | grep -v "finalize\|fold" || tail -5 gpurun_out/microbench_norm.log
timeout -k 10 900 python bench.py --steps 10 --warmup 4 --no-cpu-baseline 2>&1 | grep -v Warn | tail -1 | tee gpurun_out/bench10.json
timeout -k 10 900 python tools/timeline_step.py 16 2>&1 | grep -v Warn > gpurun_out/timeline.txt; head -45 gpurun_out/timeline.txt
