#!/bin/bash
# One gpurun call: default bench + steady-state launch list of the headline step (eager, after 3 warm-up steps).
mkdir -p gpurun_out
python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "bench exit $?"
tail -c 600 gpurun_out/bench_default.err
ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 3000 -c 900 --csv --log-file gpurun_out/launches_gan_train.csv \
  python tools/step_runner.py gen_train --B 16 --steps 3 --warmup 4 > gpurun_out/step_runner_ncu.log 2>&1
tail -2 gpurun_out/step_runner_ncu.log
