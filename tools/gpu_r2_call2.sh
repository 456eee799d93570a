#!/bin/bash
# Round 2, call 2: the promoted tests, the new headline bench (balanced step, global batch 128 -> B=128 at N=1), the B=16
# point, the reference arm, the halo A/B on the new step and a launch list of the B=16 step.
mkdir -p gpurun_out
exec > >(tee gpurun_out/r2_call2.log) 2>&1
timeout 600 python -m pytest tests -m gpu -x -q -p no:cacheprovider 2>&1 | tail -6
echo "== bench default (B=128)"
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r2_b128.json 2> gpurun_out/bench_r2_b128.err; tail -c 1500 gpurun_out/bench_r2_b128.json; tail -25 gpurun_out/bench_r2_b128.err
echo "== bench B=16"
HWG_BENCH_B=16 HWG_BENCH_NO_EXTRAS=1 timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_r2_b16.json 2> gpurun_out/bench_r2_b16.err; tail -c 1200 gpurun_out/bench_r2_b16.json; tail -5 gpurun_out/bench_r2_b16.err
echo "== reference arm"
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 | tail -c 600
for B in 16 128; do
  for mode in 0 1; do
    echo "== step_runner gan_step B=$B HWG_CONV_HALO=$mode"
    HWG_CONV_HALO=$mode timeout 300 python tools/step_runner.py gan_step --B $B --steps 10 --graph
  done
done
echo "== launch list, B=16 balanced step (eager, 2 steps)"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_gan_step_b16.csv \
  python tools/step_runner.py gan_step --B 16 --steps 2 --warmup 2 > gpurun_out/ncu_b16.log 2>&1
tail -2 gpurun_out/ncu_b16.log
python tools/parse_launches.py gpurun_out/launches_gan_step_b16.csv | head -60
