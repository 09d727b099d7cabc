mkdir -p gpurun_out
(time timeout -k 10 600 python -m pytest tests -q -m gpu -x --tb=short 2>&1 | grep -v "Warning\|warnings.html\|detach()" | tail -12) > gpurun_out/t_all.log 2>&1
tail -6 gpurun_out/t_all.log
TAG=default timeout 200 python tests/diag_graph.py 2>&1 | grep -v Warn | tail -3
