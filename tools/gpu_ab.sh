#!/bin/bash
# same-box A/B of two builds of the library: lib/libhwg_prev.so against lib/libhwg_b200.so, alternating runs
mkdir -p gpurun_out
exec > >(tee gpurun_out/ab.log) 2>&1
P=$PWD/handwriting_line_generation_b200/lib/libhwg_prev.so
for B in ${BS:-128}; do
for i in 1 2 3; do
  echo -n "B=$B prev: "; HWG_LIB_PATH=$P timeout 300 python tools/step_runner.py gan_step --B $B --steps 20 --graph 2>&1 | tail -1
  echo -n "B=$B new:  "; timeout 300 python tools/step_runner.py gan_step --B $B --steps 20 --graph 2>&1 | tail -1
done
done
