mkdir -p gpurun_out
timeout 120 python tools/bench_upsample.py 2>&1 | grep -v Warn | tee gpurun_out/bench_up_row3.log
timeout -k 10 300 python bench.py --steps 10 --warmup 4 --no-cpu-baseline 2>gpurun_out/bench29.err | tail -1 > gpurun_out/bench29.json
DWC_ROWUP=0 timeout -k 10 300 python bench.py --steps 10 --warmup 4 --no-cpu-baseline 2>gpurun_out/bench30.err | tail -1 > gpurun_out/bench30.json
cut -c1-200 gpurun_out/bench29.json gpurun_out/bench30.json; grep -o '"roofline.*' gpurun_out/bench29.json | cut -c1-400
