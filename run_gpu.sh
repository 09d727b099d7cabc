mkdir -p gpurun_out
timeout -k 10 600 python -m pytest tests/ -x -q -m gpu 2>&1 | tail -3
for b in 16 48; do echo "== merged stride-2 dgrad phases, batch $b"; timeout -k 10 600 python tools/microbench.py --batch $b --only G2,G3,G4,G5,D3,D5 2>&1 | grep "^|" | grep dgrad; done > gpurun_out/microbench_phase.log 2>&1; cat gpurun_out/microbench_phase.log
timeout -k 10 900 python bench.py --steps 10 --warmup 4 --no-cpu-baseline 2>&1 | grep -v Warn | tail -1 | tee gpurun_out/bench14.json
