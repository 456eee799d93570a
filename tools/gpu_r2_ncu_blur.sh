#!/bin/bash
# ncu --set full of the b4-level blur launches (forward with noise, backward adjoint) of one B=128 step
mkdir -p gpurun_out
exec > >(tee gpurun_out/ncu_blur.log) 2>&1
timeout 170 ncu --set full --clock-control none --import-source on -k regex:blur_noise --launch-skip 3 -c 6 -f -o /tmp/blur_r02 \
  python tools/step_runner.py gan_step --B 128 --steps 1 --warmup 0 > gpurun_out/ncu_blur_run.log 2>&1
tail -1 gpurun_out/ncu_blur_run.log
python tools/ncu_summary.py /tmp/blur_r02.ncu-rep 14 > gpurun_out/blur_r02_summary.txt 2>&1
python - <<'PY' >> gpurun_out/blur_r02_summary.txt
import csv, subprocess
raw = subprocess.run(['ncu', '-i', '/tmp/blur_r02.ncu-rep', '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
h = rows[0]
want = ['launch__occupancy_limit_shared_mem', 'launch__occupancy_limit_registers', 'launch__waves_per_multiprocessor', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio',
        'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_membar_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio', 'launch__grid_size', 'launch__shared_mem_per_block_dynamic', 'launch__occupancy_limit_blocks', 'sm__maximum_warps_per_active_cycle_pct',
        'smsp__inst_executed.sum', 'lts__t_sector_hit_rate.pct', 'dram__throughput.avg.pct_of_peak_sustained_elapsed']
for r in rows[2:]:
    print('---', r[h.index('Kernel Name')][:50], r[h.index('gpu__time_duration.sum')])
    for k in want:
        if k in h: print('   ', k, '=', r[h.index(k)])
PY
tail -40 gpurun_out/blur_r02_summary.txt
