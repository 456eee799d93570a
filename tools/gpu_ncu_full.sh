#!/bin/bash
# ncu --set full over every libhwg kernel of ONE steady-state train step; the raw metric page is exported as CSV
mkdir -p gpurun_out
ncu --set full --clock-control none -k "regex:^(adain|adam|blur|bn_|conv_|ctc_|gen_|hwr_|linear_|logsoftmax|maxpool|pixelnorm|relu_|scale_|wgrad_)" -s 692 -c 173 -o /tmp/prof_step \
  python tools/step_runner.py gen_train --B 16 --steps 2 --warmup 4 > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/ncu_full.log
ncu -i /tmp/prof_step.ncu-rep --page raw --csv > gpurun_out/prof_step_raw.csv 2>/dev/null
ls -la /tmp/prof_step.ncu-rep gpurun_out/prof_step_raw.csv
