#!/bin/bash
# same-box A/B of two builds of the library: lib/libhwg_prev.so against lib/libhwg_b200.so, alternating runs
mkdir -p gpurun_out
exec > >(tee gpurun_out/ab.log) 2>&1
P=$PWD/handwriting_line_generation_b200/lib/libhwg_prev.so
for i in 1 2 3; do
  echo -n "prev: "; HWG_LIB_PATH=$P timeout 300 python tools/step_runner.py gan_step --B ${B:-128} --steps 10 --graph 2>&1 | tail -1
  echo -n "new:  "; timeout 300 python tools/step_runner.py gan_step --B ${B:-128} --steps 10 --graph 2>&1 | tail -1
done
