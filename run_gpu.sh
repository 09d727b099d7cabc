mkdir -p gpurun_out
timeout -k 10 900 python tools/timeline_step.py 16 2>&1 | grep -v Warn > gpurun_out/timeline.txt; head -52 gpurun_out/timeline.txt
timeout -k 10 300 python bench.py --steps 10 --warmup 4 --no-cpu-baseline 2>gpurun_out/bench22.err | tail -1 > gpurun_out/bench22.json
DWC_BATCH_PACK=0 timeout -k 10 300 python bench.py --steps 10 --warmup 4 --no-cpu-baseline 2>gpurun_out/bench23.err | tail -1 > gpurun_out/bench23.json
cut -c1-200 gpurun_out/bench22.json gpurun_out/bench23.json; tail -3 gpurun_out/bench22.err
