#!/bin/bash
mkdir -p gpurun_out
exec > >(tee gpurun_out/r2_call26.log) 2>&1
run() {
  echo "== $1"
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:conv_small --launch-skip 5 -c 1 -o /tmp/s_$1 \
    python tools/conv_bench.py $1 > gpurun_out/ncu_s_$1.log 2>&1
  python tools/ncu_summary.py /tmp/s_$1.ncu-rep 24 > gpurun_out/sum_s_$1.txt 2>&1
  ncu -i /tmp/s_$1.ncu-rep --page source --csv 2>/dev/null | gzip > gpurun_out/src_s_$1.csv.gz
  cat gpurun_out/sum_s_$1.txt | cut -c1-170
}
run gen_b4c2_plain
run gen_b4c2
