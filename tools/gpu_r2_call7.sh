#!/bin/bash
# Round 2, call 7: A/B of the unit-major conv_wgrad launch order; ncu --set full of the top kernels besides conv_fprop
# (B=64 step: same regime, quicker replays), summarised as text.
mkdir -p gpurun_out
exec > >(tee gpurun_out/r2_call7.log) 2>&1
timeout 300 python -m pytest tests/test_conv_bwd_gpu.py -m gpu -q -p no:cacheprovider 2>&1 | tail -3
for ord in 0 1; do
  echo "== gan_step B=128 graph HWG_WGRAD_ORDER=$ord"
  HWG_WGRAD_ORDER=$ord timeout 300 python tools/step_runner.py gan_step --B 128 --steps 10 --graph 2>&1 | tail -1
done
for k in adain_bwd_apply wgrad_small conv_wgrad conv_small blur_noise_act norm_bwd_apply scale_shift_act gen_output_bwd relu_maxpool_bwd_win; do
  echo "== ncu --set full $k"
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k --launch-skip 2 -c 5 -o /tmp/k_$k \
    python tools/step_runner.py gan_step --B 64 --steps 1 --warmup 0 > gpurun_out/ncu_$k.log 2>&1
  tail -1 gpurun_out/ncu_$k.log
  python tools/ncu_summary.py /tmp/k_$k.ncu-rep 14 > gpurun_out/sum_$k.txt 2>&1
done
ls -la gpurun_out | tail -15
