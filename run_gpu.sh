mkdir -p gpurun_out
(time timeout -k 10 600 python -m pytest tests -q -m gpu -x --tb=short 2>&1 | grep -v "Warning\|warnings.html\|detach()" | tail -12) > gpurun_out/t_all.log 2>&1
tail -6 gpurun_out/t_all.log
timeout -k 10 300 python tools/microbench.py --batch 48 --only NONE --out gpurun_out/r01g_microbench_norm_row_b48.md > /dev/null 2>&1
timeout -k 10 300 python tools/microbench.py --batch 16 --only NONE --out gpurun_out/r01g_microbench_norm_row_b16.md > /dev/null 2>&1
timeout 120 python tools/bench_conv7.py 2>&1 | grep -v Warn > gpurun_out/r01g_bench_conv7.md
timeout -k 10 400 python bench.py --steps 20 --warmup 5 2>gpurun_out/bench28.err | tail -1 > gpurun_out/bench28.json
cat gpurun_out/bench28.json
timeout -k 10 400 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_step.csv python tools/profile_step.py 16 > gpurun_out/ncu_step.log 2>&1
python tools/summarize_launches.py gpurun_out/launches_step.csv > gpurun_out/launches_summary.txt; head -12 gpurun_out/launches_summary.txt
rm -f gpurun_out/launches_step.csv
timeout -k 10 300 ncu --set full --clock-control none --profile-from-start off -k regex:"row_kernel|conv7few" -o /tmp/prof_new -f python tools/profile_kernels.py 48 > gpurun_out/ncu_new.log 2>&1
ncu -i /tmp/prof_new.ncu-rep --page raw --csv > gpurun_out/ncu_new_raw.csv 2>/dev/null
ls -la /tmp/prof_new.ncu-rep gpurun_out/ncu_new_raw.csv; tail -2 gpurun_out/ncu_new.log
