#!/bin/bash
mkdir -p gpurun_out
exec > >(tee gpurun_out/r2_call27.log) 2>&1
timeout 900 python -m pytest tests/test_modules_gpu.py tests/test_gen_train_gpu.py tests/test_gen_bwd_ops_gpu.py tests/test_conv_gpu.py -m gpu -q -p no:cacheprovider 2>&1 | tail -5
P=$PWD/handwriting_line_generation_b200/lib/libhwg_prev.so
for i in 1 2 3; do
  echo -n "prev: "; HWG_LIB_PATH=$P timeout 300 python tools/step_runner.py gan_step --B 128 --steps 10 --graph 2>&1 | tail -1
  echo -n "new:  "; timeout 300 python tools/step_runner.py gan_step --B 128 --steps 10 --graph 2>&1 | tail -1
done
