#!/bin/bash
# Round 2, call 5: new tests (spacing kernels, CountCNN), full -m gpu, then ncu --set full over the secondary kernels of one
# B=128 step (HBM-bound passes, staged-tile conv, wgrad) to find what limits them.
mkdir -p gpurun_out
exec > >(tee gpurun_out/r2_call5.log) 2>&1
timeout 1200 python -m pytest tests -m gpu -q -p no:cacheprovider -s 2>&1 | grep -v "^$" | tail -25
echo "== ncu --set full, secondary kernels, B=128 (second step)"
K='regex:^(adain_bwd_apply|adain_bwd_reduce|scale_shift_act|conv_small|conv_wgrad|wgrad_small|blur_noise|norm_bwd_apply|act_bwd|relu_maxpool_bwd|gen_output_bwd)'
timeout 1500 ncu --set full --clock-control none --import-source on -k "$K" --launch-skip 227 -c 227 -o gpurun_out/sec_b128 \
  python tools/step_runner.py gan_step --B 128 --steps 1 --warmup 1 > gpurun_out/ncu_sec.log 2>&1
tail -2 gpurun_out/ncu_sec.log
ncu -i gpurun_out/sec_b128.ncu-rep --page raw --csv > gpurun_out/sec_b128_raw.csv 2>/dev/null
ls -la gpurun_out/sec_b128.ncu-rep gpurun_out/sec_b128_raw.csv
