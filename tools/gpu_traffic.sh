#!/bin/bash
# (1) DRAM traffic + duration of every kernel of one steady-state train step (few ncu passes per kernel)
# (2) one --set full capture of the dominant kernel (three launches) for the record
mkdir -p gpurun_out
timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none \
  --launch-skip 780 -c 200 --csv --log-file gpurun_out/traffic_gan_train.csv \
  python tools/step_runner.py gen_train --B 16 --steps 3 --warmup 4 > gpurun_out/traffic.log 2>&1
tail -1 gpurun_out/traffic.log
timeout 300 ncu --set full --clock-control none --import-source on -k "regex:conv_fprop" -s 40 -c 3 -o gpurun_out/prof_conv_fprop_r01 \
  python tools/step_runner.py gen_train --B 16 --steps 1 --warmup 1 > gpurun_out/ncu_conv.log 2>&1
tail -2 gpurun_out/ncu_conv.log
ls -la gpurun_out/*.ncu-rep gpurun_out/traffic_gan_train.csv
