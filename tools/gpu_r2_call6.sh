#!/bin/bash
# Round 2, call 6: style extractor on the GPU, the 7-lesson cycle, the full bench line, and the profile artefacts of the
# headline step (launch list + DRAM traffic per kernel; one --set full capture of a conv_fprop launch, summarised as text)
mkdir -p gpurun_out
exec > >(tee gpurun_out/r2_call6.log) 2>&1
timeout 900 python -m pytest tests/test_char_style_gpu.py tests/test_spacing_gpu.py -m gpu -q -p no:cacheprovider -s 2>&1 | grep -v "^$" | tail -15
echo "== 7-lesson cycle"
timeout 600 python bench_cycle.py --B 16 --cycles 4 2>&1 | tail -c 1500
echo "== bench default (B=128)"
timeout 1200 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r2_b128.json 2> gpurun_out/bench_r2_b128.err; tail -c 2500 gpurun_out/bench_r2_b128.json; tail -8 gpurun_out/bench_r2_b128.err
echo "== DRAM traffic + duration per kernel, one B=128 step"
timeout 1200 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none --csv \
  --log-file gpurun_out/traffic_gan_step_b128.csv python tools/step_runner.py gan_step --B 128 --steps 1 --warmup 2 > gpurun_out/traffic.log 2>&1
tail -1 gpurun_out/traffic.log
python tools/parse_traffic.py gpurun_out/traffic_gan_step_b128.csv 0 gpurun_out/traffic_gan_train_r02.json batch=128 step=balanced | tail -70
gzip -f gpurun_out/traffic_gan_step_b128.csv
echo "== ncu --set full, conv_fprop (recognizer conv3-like launches of the second step)"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:^conv_fprop_kernel --launch-skip 150 -c 6 -o /tmp/conv_fprop_r02 \
  python tools/step_runner.py gan_step --B 128 --steps 1 --warmup 1 > gpurun_out/ncu_conv.log 2>&1
tail -1 gpurun_out/ncu_conv.log
python tools/ncu_summary.py /tmp/conv_fprop_r02.ncu-rep 6 > gpurun_out/conv_fprop_r02_summary.txt 2>&1; head -c 3000 gpurun_out/conv_fprop_r02_summary.txt
ls -la gpurun_out | head -40; du -sh gpurun_out
