mkdir -p gpurun_out
timeout -k 10 200 python bench.py --steps 20 --warmup 5 2>gpurun_out/bench40.err | tail -1 > gpurun_out/bench40.json
cut -c1-230 gpurun_out/bench40.json
