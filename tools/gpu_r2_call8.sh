#!/bin/bash
# Round 2, call 8: conv_wgrad_kernel with tap groups (+ halo box of x) and the two-CTA wgrad_small variants: parity, then A/B.
mkdir -p gpurun_out
exec > >(tee gpurun_out/r2_call8.log) 2>&1
echo "== parity, defaults (halo on, wgs variant 1)"
timeout 600 python -m pytest tests/test_conv_bwd_gpu.py -m gpu -q -p no:cacheprovider -x 2>&1 | tail -15
echo "== parity, HWG_WGRAD_HALO=0 HWG_WGS_VARIANT=0"
HWG_WGRAD_HALO=0 HWG_WGS_VARIANT=0 timeout 600 python -m pytest tests/test_conv_bwd_gpu.py -m gpu -q -p no:cacheprovider 2>&1 | tail -8
for cfg in "0 0" "0 1" "1 0" "1 1"; do
  set -- $cfg
  echo "== gan_step B=128 graph HWG_WGRAD_HALO=$1 HWG_WGS_VARIANT=$2"
  HWG_WGRAD_HALO=$1 HWG_WGS_VARIANT=$2 timeout 300 python tools/step_runner.py gan_step --B 128 --steps 10 --graph 2>&1 | tail -1
done
for cfg in "0 0" "1 1"; do
  set -- $cfg
  echo "== gan_step B=16 graph HWG_WGRAD_HALO=$1 HWG_WGS_VARIANT=$2"
  HWG_WGRAD_HALO=$1 HWG_WGS_VARIANT=$2 timeout 300 python tools/step_runner.py gan_step --B 16 --steps 20 --graph 2>&1 | tail -1
done
echo "== module-level parity with the new kernels"
timeout 900 python -m pytest tests/test_gen_train_gpu.py tests/test_hwr_train_gpu.py tests/test_disc_gpu.py tests/test_char_style_gpu.py tests/test_spacing_gpu.py -m gpu -q -p no:cacheprovider 2>&1 | tail -8
