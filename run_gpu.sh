mkdir -p gpurun_out
python dwc_gan_b200/build.py > gpurun_out/build.log 2>&1 || tail -5 gpurun_out/build.log
for f in test_modules_gpu test_step_gpu; do
  timeout -k 10 1200 python -m pytest tests/$f.py -q -m gpu -x --tb=short 2>&1 | tail -60 > gpurun_out/$f.log
  echo "== $f"; tail -45 gpurun_out/$f.log
done
