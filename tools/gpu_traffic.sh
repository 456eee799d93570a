#!/bin/bash
# DRAM traffic + duration of every kernel of steady-state train steps (few ncu passes per kernel); tools/parse_traffic.py
# cuts ONE step out of the capture (between two adam_flat_kernel launches)
mkdir -p gpurun_out
timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none \
  --launch-skip ${SKIP:-1100} -c ${COUNT:-560} --csv --log-file gpurun_out/traffic_gan_train.csv \
  python tools/step_runner.py gen_train --B 16 --steps 3 --warmup 4 > gpurun_out/traffic.log 2>&1
tail -1 gpurun_out/traffic.log
python tools/parse_traffic.py gpurun_out/traffic_gan_train.csv 229 gpurun_out/traffic_gan_train.json | tail -45
