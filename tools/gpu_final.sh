#!/bin/bash
# round-end evidence on one GPU (artifacts kept small: gpurun_out/ is capped at 64 MiB): per-kernel traffic pass of one
# steady-state step, the default bench line + reference arm, --set full summary of three discriminator conv launches
mkdir -p gpurun_out
bash tools/gpu_traffic.sh | tail -2
cp gpurun_out/traffic_gan_train.json profiles/traffic_gan_train_r01.json      # bench.py reads roofline.traffic from here
timeout 900 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "bench exit $?"
timeout 300 python bench.py --impl reference --steps 6 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; echo "reference exit $?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_default.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], 'launches/step', d['gpu_launches']/d['steps'], 'traffic', d['roofline']['traffic'])
PY
timeout 300 ncu --set full --clock-control none --import-source on -k "regex:conv_fprop" -s 80 -c 3 -o /tmp/prof_conv_disc \
  python tools/step_runner.py gen_train --B 16 --steps 1 --warmup 1 > gpurun_out/ncu_conv.log 2>&1
tail -1 gpurun_out/ncu_conv.log
python tools/ncu_summary.py /tmp/prof_conv_disc.ncu-rep 14 > gpurun_out/conv_fprop_disc_r01_summary.txt 2>&1
head -48 gpurun_out/conv_fprop_disc_r01_summary.txt
du -sh gpurun_out
