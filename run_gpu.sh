mkdir -p gpurun_out
timeout -k 10 400 python bench.py --steps 20 --warmup 5 2>gpurun_out/bench37.err | tail -1 > gpurun_out/bench37.json
cut -c1-260 gpurun_out/bench37.json
