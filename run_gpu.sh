mkdir -p gpurun_out
python dwc_gan_b200/build.py > gpurun_out/build.log 2>&1 || tail -5 gpurun_out/build.log
timeout -k 10 900 python -m pytest tests/test_conv_gpu.py tests/test_rows_gpu.py tests/test_post_gpu.py -q -m gpu -n 4 --tb=short 2>&1 | tail -40 > gpurun_out/t_units.log; tail -40 gpurun_out/t_units.log
timeout -k 10 1500 python -m pytest tests/test_modules_gpu.py tests/test_step_gpu.py -q -m gpu -n 4 --tb=short -rP 2>&1 | grep -v "^frame\|python()\|Warning\|warnings.html\|detach()" | tail -60 > gpurun_out/t_net.log; tail -60 gpurun_out/t_net.log
