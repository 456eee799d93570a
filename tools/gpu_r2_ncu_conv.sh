#!/bin/bash
# ncu --set full of every conv_fprop_kernel launch of ONE B=128 step with the final library (tensor-pipe activity per launch),
# after one pass of the GPU suite
mkdir -p gpurun_out
exec > >(tee gpurun_out/ncu_conv_final.log) 2>&1
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider 2>&1 | grep -E "^E  .*(Assertion|assert)|^FAILED|passed|failed" | head -12
SKIP=${SKIP:-85}; CNT=${CNT:-85}
timeout ${NCU_TIMEOUT:-480} ncu --set full --clock-control none --import-source on -k regex:^conv_fprop_kernel --launch-skip $SKIP -c $CNT -f -o /tmp/conv_fprop_r02f \
  python tools/step_runner.py gan_step --B 128 --steps 1 --warmup 1 > gpurun_out/ncu_conv_f.log 2>&1
tail -1 gpurun_out/ncu_conv_f.log
python tools/ncu_summary.py /tmp/conv_fprop_r02f.ncu-rep 8 > gpurun_out/conv_fprop_r02_final_summary.txt 2>&1
python - <<'PY' > gpurun_out/conv_fprop_r02_final_tensor_pipe.txt
import csv, subprocess
raw = subprocess.run(['ncu', '-i', '/tmp/conv_fprop_r02f.ncu-rep', '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
h = rows[0]
def col(k): return h.index(k)
keys = ['gpu__time_duration.sum', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'launch__grid_size']
print('conv_fprop launches of one B=128 step, final library: duration_us tensor_pipe_active_pct dram_read dram_write grid | kernel')
tot = 0.0; wsum = 0.0
for r in rows[2:]:
    d, tp = float(r[col(keys[0])]), float(r[col(keys[1])])
    tot += d; wsum += d * tp
    print(f"{d:9.1f} {tp:6.1f} {r[col(keys[2])][:8]:>9} {r[col(keys[3])][:8]:>9} {r[col(keys[4])]:>5} | {r[col('Kernel Name')][:60]}")
print(f"time-weighted tensor-pipe activity over these {len(rows) - 2} launches: {wsum / tot:.1f} %  ({tot:.0f} us)")
PY
tail -3 gpurun_out/conv_fprop_r02_final_tensor_pipe.txt
