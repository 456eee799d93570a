#!/bin/bash
# Round 2 evidence of the headline step with the final kernels: DRAM traffic + duration per kernel of one B=128 step,
# launch list, ncu --set full of the conv_fprop launches (tensor-pipe activity), and of one launch of each other kernel family.
mkdir -p gpurun_out
exec > >(tee gpurun_out/r2_evidence.log) 2>&1
echo "== DRAM traffic + duration per kernel, one B=128 step"
timeout 1200 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none --csv \
  --log-file gpurun_out/traffic_gan_step_b128.csv python tools/step_runner.py gan_step --B 128 --steps 1 --warmup 2 > gpurun_out/traffic.log 2>&1
tail -1 gpurun_out/traffic.log
python tools/parse_traffic.py gpurun_out/traffic_gan_step_b128.csv 0 gpurun_out/traffic_gan_train_r02.json batch=128 step=balanced | tail -75
gzip -f gpurun_out/traffic_gan_step_b128.csv
echo "== ncu --set full: conv_fprop launches of one B=128 step (first 60 of the second step)"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:^conv_fprop_kernel --launch-skip 92 -c 60 -o /tmp/conv_fprop_r02 \
  python tools/step_runner.py gan_step --B 128 --steps 1 --warmup 1 > gpurun_out/ncu_conv.log 2>&1
tail -1 gpurun_out/ncu_conv.log
python tools/ncu_summary.py /tmp/conv_fprop_r02.ncu-rep 8 > gpurun_out/conv_fprop_r02_summary.txt 2>&1
python - <<'PY'
import csv, subprocess
raw = subprocess.run(['ncu', '-i', '/tmp/conv_fprop_r02.ncu-rep', '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
h = rows[0]
def col(k): return h.index(k)
keys = ['gpu__time_duration.sum', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'launch__grid_size']
print('conv_fprop launches: duration_us tensor_pipe_active_pct dram_read_MB dram_write_MB grid | kernel')
tot = 0.0; wsum = 0.0
for r in rows[2:]:
    d, tp = float(r[col(keys[0])]), float(r[col(keys[1])])
    tot += d; wsum += d * tp
    print(f"{d:9.1f} {tp:6.1f} {r[col(keys[2])][:8]:>9} {r[col(keys[3])][:8]:>9} {r[col(keys[4])]:>5} | {r[col('Kernel Name')][:60]}")
print(f"time-weighted tensor-pipe activity over these launches: {wsum / tot:.1f} %  ({tot:.0f} us)")
PY
