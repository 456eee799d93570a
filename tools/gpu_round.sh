#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"
tail -15 gpurun_out/pytest_gpu.log
HWG_BENCH_NO_EXTRAS=1 python bench.py --steps 30 > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "bench exit $?"
tail -c 600 gpurun_out/bench_default.err
cut -c1-250 gpurun_out/bench_default.json
ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 2400 -c 700 --csv --log-file gpurun_out/launches_gan_train.csv \
  python tools/step_runner.py gen_train --B 16 --steps 3 --warmup 4 > gpurun_out/step_runner_ncu.log 2>&1
tail -2 gpurun_out/step_runner_ncu.log
