#!/bin/bash
mkdir -p gpurun_out
exec > >(tee gpurun_out/r2_call20.log) 2>&1
timeout 900 python -m pytest tests/test_conv_gpu.py tests/test_disc_gpu.py tests/test_enc_gpu.py -m gpu -q -p no:cacheprovider 2>&1 | tail -3
timeout 300 python tools/step_runner.py gan_step --B 128 --steps 10 --graph 2>&1 | tail -1
timeout 300 python tools/step_runner.py gan_step --B 16 --steps 20 --graph 2>&1 | tail -1
