#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"
tail -5 gpurun_out/pytest_gpu.log
python tools/step_runner.py gen_train --B 16 --steps 20 --graph 2>&1 | tail -1
ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 780 -c 200 --csv --log-file gpurun_out/launches_gan_train.csv \
  python tools/step_runner.py gen_train --B 16 --steps 3 --warmup 4 > gpurun_out/step_runner_ncu.log 2>&1
