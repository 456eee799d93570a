#!/bin/bash
# ncu --set full on single conv_fprop launches (conv_bench shapes): the 64->64 tall discriminator layer with and without
# the halo loop, the in_conv 7-tap layer, generator mid blocks; raw metric dump for offline reading.
mkdir -p gpurun_out
exec > >(tee gpurun_out/r2_call10.log) 2>&1
run() {  # name env shape
  echo "== $1"
  env $2 timeout 300 ncu --set full --clock-control none --import-source on -k regex:conv_fprop --launch-skip 5 -c 1 -o /tmp/c_$1 \
    python tools/conv_bench.py $3 > gpurun_out/ncu_c_$1.log 2>&1
  tail -2 gpurun_out/ncu_c_$1.log
  python tools/ncu_summary.py /tmp/c_$1.ncu-rep 16 > gpurun_out/sum_c_$1.txt 2>&1
  ncu -i /tmp/c_$1.ncu-rep --page raw --csv > gpurun_out/raw_c_$1.csv 2>/dev/null
  ncu -i /tmp/c_$1.ncu-rep --page source --csv 2>/dev/null | gzip > gpurun_out/src_c_$1.csv.gz
}
run disc10_halo "HWG_CONV_TILE_W=32" t_disc_convs1_0
run disc10_nohalo "HWG_CONV_TILE_W=32 HWG_CONV_HALO=0" t_disc_convs1_0
run disc20_halo "HWG_CONV_TILE_W=32" t_disc_convs2_0
run gen_b1c2 "A=1" t_gen_b1c2
run gen_b2c2 "A=1" t_gen_b2c2
run hwr_conv5 "A=1" t_hwr_conv5
echo "== plain timings"
HWG_CONV_TILE_W=32 python tools/conv_bench.py t_disc_convs1_0 t_disc_convs1_3 t_disc_convs2_0 t_gen_b1c2 t_gen_b2c2 t_hwr_conv5 t_hwr_conv1
ls -la gpurun_out | tail -25
