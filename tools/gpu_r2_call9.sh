#!/bin/bash
mkdir -p gpurun_out
exec > >(tee gpurun_out/r2_call9.log) 2>&1
for cfg in "0 0" "1 1"; do
  set -- $cfg
  echo "== wgrad_bench B=128 HWG_WGRAD_HALO=$1 HWG_WGS_VARIANT=$2"
  HWG_WGRAD_HALO=$1 HWG_WGS_VARIANT=$2 timeout 300 python tools/wgrad_bench.py --B 128 2>&1 | tail -16
done
echo "== wgrad_bench B=16 defaults"
timeout 300 python tools/wgrad_bench.py --B 16 2>&1 | tail -16
echo "== HWG_WGRAD_CTAS=148 (halo)"
HWG_WGRAD_CTAS=148 timeout 300 python tools/wgrad_bench.py --B 128 b2c2 b1c2 b3c1 hwr disc 2>&1 | tail -8
