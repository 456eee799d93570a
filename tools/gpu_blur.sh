#!/bin/bash
mkdir -p gpurun_out
exec > >(tee gpurun_out/blur.log) 2>&1
timeout 600 python -m pytest tests/test_bwd_ops_gpu.py tests/test_hwr_gpu.py tests/test_hwr_train_gpu.py -q -m gpu -x 2>&1 | tail -5
BS="128" bash tools/gpu_ab.sh
