#!/bin/bash
# Round 2, call 3: after breaking the output -> grad_fn -> ctx cycle (graph capture of the balanced step), BASELINE-size parity
mkdir -p gpurun_out
exec > >(tee gpurun_out/r2_call3.log) 2>&1
timeout 900 python -m pytest tests -m gpu -x -q -p no:cacheprovider 2>&1 | tail -8
echo "== bench default (B=128)"
HWG_BENCH_DUMP_CONV=gpurun_out/conv_b128.json timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r2_b128.json 2> gpurun_out/bench_r2_b128.err; tail -c 2500 gpurun_out/bench_r2_b128.json; tail -12 gpurun_out/bench_r2_b128.err
python tools/conv_table.py gpurun_out/conv_b128.json | head -70
echo "== bench B=16"
HWG_BENCH_B=16 HWG_BENCH_NO_EXTRAS=1 HWG_BENCH_NO_CPU_BASELINE=1 HWG_BENCH_DUMP_CONV=gpurun_out/conv_b16.json timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_r2_b16.json 2> gpurun_out/bench_r2_b16.err; tail -c 800 gpurun_out/bench_r2_b16.json; tail -5 gpurun_out/bench_r2_b16.err
for B in 16 128; do
  for mode in 0 1; do
    echo "== step_runner gan_step B=$B HWG_CONV_HALO=$mode"
    HWG_CONV_HALO=$mode timeout 300 python tools/step_runner.py gan_step --B $B --steps 10 --graph 2>&1 | tail -3
  done
done
echo "== launch list, B=128 balanced step (eager, 1 step after 2 warm-ups)"
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_gan_step_b128.csv \
  python tools/step_runner.py gan_step --B 128 --steps 1 --warmup 2 > gpurun_out/ncu_b128.log 2>&1
tail -2 gpurun_out/ncu_b128.log
python tools/parse_launches.py gpurun_out/launches_gan_step_b128.csv > gpurun_out/launches_gan_step_b128.txt; head -70 gpurun_out/launches_gan_step_b128.txt
