#!/bin/bash
mkdir -p gpurun_out
exec > >(tee gpurun_out/tmp.log) 2>&1
timeout 900 python -m pytest tests/test_disc_gpu.py tests/test_enc_gpu.py tests/test_char_style_gpu.py tests/test_spacing_gpu.py tests/test_trainer_gen_gpu.py -m gpu -q -p no:cacheprovider 2>&1 | tail -4
BS="128 16" bash tools/gpu_ab.sh
