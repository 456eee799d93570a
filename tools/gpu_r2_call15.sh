#!/bin/bash
mkdir -p gpurun_out
exec > >(tee gpurun_out/r2_call15.log) 2>&1
echo "== parity (all conv / module tests)"
timeout 900 python -m pytest tests/test_conv_gpu.py tests/test_conv_bwd_gpu.py tests/test_modules_gpu.py tests/test_disc_gpu.py tests/test_enc_gpu.py -m gpu -q -p no:cacheprovider -x 2>&1 | tail -4
SH="t_disc_convs1_0 t_disc_convs1_3 t_disc_convs2_0 t_hwr_conv1 t_hwr_conv2 t_hwr_conv5 t_gen_b2c2"
echo "== conv_bench"
HWG_CONV_TILE_W=32 timeout 300 python tools/conv_bench.py $SH
echo "== gan_step"
timeout 300 python tools/step_runner.py gan_step --B 128 --steps 10 --graph 2>&1 | tail -1
timeout 300 python tools/step_runner.py gan_step --B 16 --steps 20 --graph 2>&1 | tail -1
echo "== bench B=128 (per-layer dump)"
HWG_BENCH_NO_EXTRAS=1 HWG_BENCH_NO_CPU_BASELINE=1 HWG_BENCH_NO_GPU_BASELINE=1 HWG_BENCH_DUMP_CONV=gpurun_out/conv_b128_v2.json timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_v2_b128.json 2> gpurun_out/bench_v2_b128.err; tail -c 1500 gpurun_out/bench_v2_b128.json; tail -5 gpurun_out/bench_v2_b128.err
