#!/bin/bash
mkdir -p gpurun_out
exec > >(tee gpurun_out/r2_call17.log) 2>&1
for i in 1 2 3; do
  timeout 300 python -m pytest tests/test_baseline_sizes_gpu.py -m gpu -q -p no:cacheprovider -k gradient_sets -s 2>&1 | grep -E "AssertionError:|passed|failed|worst per-tensor"
done
echo "== HWG_WGS_VARIANT=0"
HWG_WGS_VARIANT=0 timeout 300 python -m pytest tests/test_baseline_sizes_gpu.py -m gpu -q -p no:cacheprovider -k gradient_sets -s 2>&1 | grep -E "AssertionError:|passed|failed|worst per-tensor"
echo "== HWG_CONV_HALO=0"
HWG_CONV_HALO=0 timeout 300 python -m pytest tests/test_baseline_sizes_gpu.py -m gpu -q -p no:cacheprovider -k gradient_sets -s 2>&1 | grep -E "AssertionError:|passed|failed|worst per-tensor"
echo "== wgrad parity + bench (new wgrad_small loop)"
timeout 300 python -m pytest tests/test_conv_bwd_gpu.py -m gpu -q -p no:cacheprovider 2>&1 | tail -2
timeout 300 python tools/wgrad_bench.py --B 128 b4c2 b4c1 b3c2 2>&1 | tail -4
HWG_WGS_VARIANT=0 timeout 300 python tools/wgrad_bench.py --B 128 b4c2 b4c1 b3c2 2>&1 | tail -4
timeout 300 python tools/step_runner.py gan_step --B 128 --steps 10 --graph 2>&1 | tail -1
