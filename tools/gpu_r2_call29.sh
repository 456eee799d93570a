#!/bin/bash
mkdir -p gpurun_out
exec > >(tee gpurun_out/r2_call29.log) 2>&1
timeout 900 python -m pytest tests/test_enc_gpu.py tests/test_trainer_gen_gpu.py tests/test_balance_gpu.py -m gpu -q -p no:cacheprovider 2>&1 | tail -12
echo "== trainer_gen with the old route"
HWG_NO_STEM_CONV=1 timeout 900 python -m pytest tests/test_trainer_gen_gpu.py -m gpu -q -p no:cacheprovider 2>&1 | tail -4
for i in 1 2 3; do
  echo -n "old route: "; HWG_NO_STEM_CONV=1 timeout 300 python tools/step_runner.py gan_step --B 128 --steps 10 --graph 2>&1 | tail -1
  echo -n "stem_conv: "; timeout 300 python tools/step_runner.py gan_step --B 128 --steps 10 --graph 2>&1 | tail -1
done
echo -n "B=16 old route: "; HWG_NO_STEM_CONV=1 timeout 300 python tools/step_runner.py gan_step --B 16 --steps 20 --graph 2>&1 | tail -1
echo -n "B=16 stem_conv: "; timeout 300 python tools/step_runner.py gan_step --B 16 --steps 20 --graph 2>&1 | tail -1
