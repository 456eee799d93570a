#!/bin/bash
mkdir -p gpurun_out
exec > >(tee gpurun_out/r2_call25.log) 2>&1
timeout 600 python -m pytest tests/test_conv_gpu.py tests/test_modules_gpu.py tests/test_gen_train_gpu.py -m gpu -q -p no:cacheprovider 2>&1 | tail -2
for e in 2 3; do echo "== conv_bench HWG_CS_MINB=$e"; HWG_CS_MINB=$e python tools/conv_bench.py gen_b4c2 gen_b4c2_plain gen_b3c2; done
for i in 1 2 3; do
  echo -n "MINB=2: "; HWG_CS_MINB=2 timeout 300 python tools/step_runner.py gan_step --B 128 --steps 10 --graph 2>&1 | tail -1
  echo -n "MINB=3: "; timeout 300 python tools/step_runner.py gan_step --B 128 --steps 10 --graph 2>&1 | tail -1
done
