#!/bin/bash
# Round 2, final evidence part 2: DRAM traffic per kernel of one B=128 step with the final library, then the default bench
# line (which reads the traffic summary) and the 16-line bench line.
mkdir -p gpurun_out
exec > >(tee gpurun_out/r2_final2.log) 2>&1
echo "== DRAM traffic + duration per kernel, one B=128 step"
timeout 1200 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none --csv \
  --log-file gpurun_out/traffic_gan_step_b128.csv python tools/step_runner.py gan_step --B 128 --steps 1 --warmup 2 > gpurun_out/traffic.log 2>&1
tail -1 gpurun_out/traffic.log
python tools/parse_traffic.py gpurun_out/traffic_gan_step_b128.csv 0 gpurun_out/traffic_gan_train_r02.json batch=128 step=balanced | tail -60
cp gpurun_out/traffic_gan_train_r02.json profiles/traffic_gan_train_r02.json
gzip -f gpurun_out/traffic_gan_step_b128.csv
echo "== bench default (B=128)"
timeout 1500 python bench.py > gpurun_out/bench_r02_final_b128.json 2> gpurun_out/bench_r02_final_b128.err; tail -c 300 gpurun_out/bench_r02_final_b128.json; tail -3 gpurun_out/bench_r02_final_b128.err
echo "== bench B=16"
HWG_BENCH_B=16 HWG_BENCH_NO_EXTRAS=1 HWG_BENCH_NO_CPU_BASELINE=1 HWG_BENCH_NO_GPU_BASELINE=1 timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r02_final_b16.json 2> gpurun_out/bench_r02_final_b16.err; tail -c 300 gpurun_out/bench_r02_final_b16.json
