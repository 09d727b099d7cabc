mkdir -p gpurun_out
timeout -k 10 900 python tools/timeline_step.py 16 2>&1 | grep -v Warn > gpurun_out/timeline.txt; head -60 gpurun_out/timeline.txt
