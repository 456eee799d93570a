#!/bin/bash
mkdir -p gpurun_out
exec > >(tee gpurun_out/r2_call19.log) 2>&1
timeout 900 python -m pytest tests/test_gen_bwd_ops_gpu.py tests/test_bwd_ops_gpu.py tests/test_gen_train_gpu.py tests/test_hwr_train_gpu.py -m gpu -q -p no:cacheprovider 2>&1 | tail -4
timeout 300 python tools/step_runner.py gan_step --B 128 --steps 10 --graph 2>&1 | tail -1
timeout 300 python tools/step_runner.py gan_step --B 16 --steps 20 --graph 2>&1 | tail -1
