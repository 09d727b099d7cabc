mkdir -p gpurun_out
timeout -k 10 400 python bench.py --steps 20 --warmup 5 2>gpurun_out/bench34.err | tail -1 > gpurun_out/bench34.json
cat gpurun_out/bench34.json | cut -c1-250
timeout -k 10 300 python tools/timeline_step.py 16 2>&1 | grep -v Warn > gpurun_out/timeline.txt; head -10 gpurun_out/timeline.txt
timeout -k 10 400 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_step.csv python tools/profile_step.py 16 > gpurun_out/ncu_step.log 2>&1
python tools/summarize_launches.py gpurun_out/launches_step.csv > gpurun_out/launches_summary.txt; head -3 gpurun_out/launches_summary.txt
rm -f gpurun_out/launches_step.csv
timeout -k 10 200 python tools/microbench.py --batch 48 --only G2,G3,G7,G8,G9,D2,D3,D4 --out gpurun_out/r01h_microbench_conv_b48.md > /dev/null 2>&1
timeout -k 10 200 python tools/microbench.py --batch 16 --only G2,G3,G7,G8,G9,D2,D3,D4 --out gpurun_out/r01h_microbench_conv_b16.md > /dev/null 2>&1
tail -5 gpurun_out/r01h_microbench_conv_b48.md
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
