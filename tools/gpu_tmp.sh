#!/bin/bash
mkdir -p gpurun_out
exec > >(tee gpurun_out/tmp.log) 2>&1
timeout 900 python -m pytest tests/test_bwd_ops_gpu.py tests/test_hwr_train_gpu.py tests/test_modules_gpu.py tests/test_disc_gpu.py tests/test_enc_gpu.py tests/test_baseline_sizes_gpu.py -m gpu -q -p no:cacheprovider 2>&1 | tail -4
BS="128" bash tools/gpu_ab.sh
