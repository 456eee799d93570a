#!/bin/bash
S="t_hwr_1d t_hwr_conv6 t_hwr_conv5 t_hwr_conv4 t_hwr_conv2 t_hwr_conv1 t_gen_b0c2 t_gen_b0c2_dgrad t_gen_b1c2 t_gen_b2c2"
for bn in default 256 128 64; do
  echo "== BN $bn"
  if [ $bn = default ]; then python tools/conv_bench.py $S; else HWG_CONV_BN=$bn python tools/conv_bench.py $S; fi
done
python -m pytest tests/test_optim_gpu.py -x -q 2>&1 | tail -3
