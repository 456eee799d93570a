#!/bin/bash
# Round 2, call 4: all GPU tests (halo loop now default; BASELINE-size parity; gradient sets at line size), the three-branch
# step with stash-slot sinks at B=128 and B=16, A/B against the serial issue order.
mkdir -p gpurun_out
exec > >(tee gpurun_out/r2_call4.log) 2>&1
timeout 1200 python -m pytest tests -m gpu -q -p no:cacheprovider -s 2>&1 | grep -v "^$" | tail -40
echo "== bench default (B=128)"
HWG_BENCH_DUMP_CONV=gpurun_out/conv_b128.json timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r2_b128.json 2> gpurun_out/bench_r2_b128.err; tail -c 1800 gpurun_out/bench_r2_b128.json; tail -12 gpurun_out/bench_r2_b128.err
echo "== bench B=16"
HWG_BENCH_B=16 HWG_BENCH_NO_EXTRAS=1 HWG_BENCH_NO_CPU_BASELINE=1 HWG_BENCH_NO_GPU_BASELINE=1 timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_r2_b16.json 2> gpurun_out/bench_r2_b16.err; tail -c 700 gpurun_out/bench_r2_b16.json; tail -5 gpurun_out/bench_r2_b16.err
for B in 16 32 64 128; do
  for ov in "" 1; do
    echo "== step_runner gan_step B=$B HWG_BENCH_NO_OVERLAP=$ov"
    HWG_BENCH_NO_OVERLAP=$ov timeout 300 python tools/step_runner.py gan_step --B $B --steps 10 --graph 2>&1 | tail -2
  done
done
