#!/bin/bash
mkdir -p gpurun_out
exec > >(tee gpurun_out/r2_call21.log) 2>&1
timeout 900 python -m pytest tests/test_conv_gpu.py tests/test_modules_gpu.py tests/test_disc_gpu.py tests/test_enc_gpu.py tests/test_hwr_train_gpu.py tests/test_gen_train_gpu.py -m gpu -q -p no:cacheprovider 2>&1 | tail -3
B=128 bash tools/gpu_ab.sh
