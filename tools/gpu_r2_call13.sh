#!/bin/bash
mkdir -p gpurun_out
exec > >(tee gpurun_out/r2_call13.log) 2>&1
run() {  # name env shape
  echo "== $1"
  env $2 timeout 300 ncu --set full --clock-control none --import-source on -k regex:conv_fprop --launch-skip 5 -c 1 -o /tmp/c_$1 \
    python tools/conv_bench.py $3 > gpurun_out/ncu_c_$1.log 2>&1
  tail -1 gpurun_out/ncu_c_$1.log
  python tools/ncu_summary.py /tmp/c_$1.ncu-rep 10 > gpurun_out/sum_c_$1.txt 2>&1
  ncu -i /tmp/c_$1.ncu-rep --page raw --csv > gpurun_out/raw_c_$1.csv 2>/dev/null
  ncu -i /tmp/c_$1.ncu-rep --page source --csv 2>/dev/null | gzip > gpurun_out/src_c_$1.csv.gz
}
run disc10_v2 "HWG_CONV_TILE_W=32" t_disc_convs1_0
run gen_b1c2_v2 "A=1" t_gen_b1c2
run gen_b2c2_v2 "A=1" t_gen_b2c2
run hwr_conv6_v2 "A=1" t_hwr_conv6
