timeout 40 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
DWC_CG2=2 timeout 30 python -m pytest tests/test_conv_gpu.py -q -x -k "test_conv_fwd_dgrad_wgrad and tc" --tb=line 2>&1 | tail -4
