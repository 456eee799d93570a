#!/bin/bash
mkdir -p gpurun_out
exec > >(tee gpurun_out/r2_call24.log) 2>&1
timeout 300 python -m pytest tests/test_conv_bwd_gpu.py -m gpu -q -p no:cacheprovider 2>&1 | tail -2
for i in 1 2; do
  for B in 16 128; do
    echo -n "B=$B CTAS=296: "; HWG_WGRAD_CTAS=296 timeout 300 python tools/step_runner.py gan_step --B $B --steps 20 --graph 2>&1 | tail -1
    echo -n "B=$B adaptive: "; timeout 300 python tools/step_runner.py gan_step --B $B --steps 20 --graph 2>&1 | tail -1
  done
done
for B in 32 64; do
  echo -n "B=$B CTAS=296: "; HWG_WGRAD_CTAS=296 timeout 300 python tools/step_runner.py gan_step --B $B --steps 20 --graph 2>&1 | tail -1
  echo -n "B=$B adaptive: "; timeout 300 python tools/step_runner.py gan_step --B $B --steps 20 --graph 2>&1 | tail -1
done
