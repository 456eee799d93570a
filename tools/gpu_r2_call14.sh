#!/bin/bash
mkdir -p gpurun_out
exec > >(tee gpurun_out/r2_call14.log) 2>&1
echo "== parity"
timeout 900 python -m pytest tests/test_conv_gpu.py tests/test_conv_bwd_gpu.py tests/test_modules_gpu.py tests/test_disc_gpu.py tests/test_enc_gpu.py -m gpu -q -p no:cacheprovider -x 2>&1 | tail -8
SH="t_disc_convs1_0 t_disc_convs1_3 t_disc_convs2_0 t_disc_convs3_0 t_gen_b0c2 t_gen_b1c2 t_gen_b2c2 t_hwr_conv1 t_hwr_conv2 t_hwr_conv5 t_hwr_conv6 t_hwr_1d t_gen_b0c2_dgrad"
echo "== conv_bench"
HWG_CONV_TILE_W=32 timeout 300 python tools/conv_bench.py $SH
echo "== gan_step"
timeout 300 python tools/step_runner.py gan_step --B 128 --steps 10 --graph 2>&1 | tail -1
timeout 300 python tools/step_runner.py gan_step --B 16 --steps 20 --graph 2>&1 | tail -1
