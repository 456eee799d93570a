#!/bin/bash
mkdir -p gpurun_out
exec > >(tee gpurun_out/r2_call18.log) 2>&1
echo "== launch list B=16 (ncu durations, one step)"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_gan_step_b16_v2.csv \
  python tools/step_runner.py gan_step --B 16 --steps 1 --warmup 2 > gpurun_out/ncu_b16_v2.log 2>&1
tail -1 gpurun_out/ncu_b16_v2.log
python tools/parse_launches.py gpurun_out/launches_gan_step_b16_v2.csv 2>/dev/null | head -70
gzip -f gpurun_out/launches_gan_step_b16_v2.csv
