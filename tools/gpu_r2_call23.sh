#!/bin/bash
mkdir -p gpurun_out
exec > >(tee gpurun_out/r2_call23.log) 2>&1
for c in 296 148 74; do
  echo "== wgrad_bench B=16 HWG_WGRAD_CTAS=$c"
  HWG_WGRAD_CTAS=$c timeout 300 python tools/wgrad_bench.py --B 16 b3c1 b2c2 b2c1 b1c2 b1c1 b0c2 2>&1 | tail -6
done
for c in 296 148; do
  echo "== gan_step B=16 HWG_WGRAD_CTAS=$c"; HWG_WGRAD_CTAS=$c timeout 300 python tools/step_runner.py gan_step --B 16 --steps 20 --graph 2>&1 | tail -1
done
