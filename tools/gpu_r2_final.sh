#!/bin/bash
# Round 2, final evidence on one B200: the whole GPU suite, smoke(), the default bench line (B=128, with the stock-PyTorch
# and CPU baselines and the extra workloads), the reference arm, the 16-line bench line, a launch list of one step.
mkdir -p gpurun_out
exec > >(tee gpurun_out/r2_final.log) 2>&1
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider 2>&1 | grep -v "^$" | tail -6
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
echo "== bench default (B=128)"
timeout 1500 python bench.py > gpurun_out/bench_r02_final_b128.json 2> gpurun_out/bench_r02_final_b128.err; tail -c 600 gpurun_out/bench_r02_final_b128.json; tail -3 gpurun_out/bench_r02_final_b128.err
echo "== bench --impl reference"
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_r02_final_reference.json 2> gpurun_out/bench_r02_final_reference.err; tail -c 700 gpurun_out/bench_r02_final_reference.json
echo "== bench B=16"
HWG_BENCH_B=16 HWG_BENCH_NO_EXTRAS=1 HWG_BENCH_NO_CPU_BASELINE=1 HWG_BENCH_NO_GPU_BASELINE=1 timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r02_final_b16.json 2> gpurun_out/bench_r02_final_b16.err; tail -c 400 gpurun_out/bench_r02_final_b16.json
echo "== launch list, one B=128 step"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_gan_step_b128_final.csv \
  python tools/step_runner.py gan_step --B 128 --steps 1 --warmup 2 > gpurun_out/ncu_b128_final.log 2>&1
python tools/parse_launches.py gpurun_out/launches_gan_step_b128_final.csv > gpurun_out/launches_gan_step_b128_final.txt; head -30 gpurun_out/launches_gan_step_b128_final.txt; tail -1 gpurun_out/launches_gan_step_b128_final.txt
gzip -f gpurun_out/launches_gan_step_b128_final.csv
