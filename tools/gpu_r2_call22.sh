#!/bin/bash
mkdir -p gpurun_out
exec > >(tee gpurun_out/r2_call22.log) 2>&1
P=$PWD/handwriting_line_generation_b200/lib/libhwg_prev.so
SH="t_disc_in_conv t_disc_in_conv_plain t_disc_in_conv_dgrad t_gen_b1c2 t_gen_b2c2 t_hwr_conv2 t_hwr_conv4 t_disc_convs3_0"
echo "== prev"; HWG_LIB_PATH=$P HWG_CONV_TILE_W=32 python tools/conv_bench.py $SH
echo "== new";  HWG_CONV_TILE_W=32 python tools/conv_bench.py $SH
