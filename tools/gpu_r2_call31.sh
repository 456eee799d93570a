#!/bin/bash
mkdir -p gpurun_out
exec > >(tee gpurun_out/r2_call31.log) 2>&1
for i in 1 2 3; do timeout 300 python -m pytest tests/test_trainer_gen_gpu.py -m gpu -q -p no:cacheprovider 2>&1 | grep -E "^E  |passed|failed" | head -5; done
echo "== old route"
for i in 1 2 3; do HWG_NO_STEM_CONV=1 timeout 300 python -m pytest tests/test_trainer_gen_gpu.py -m gpu -q -p no:cacheprovider 2>&1 | grep -E "^E  |passed|failed" | head -5; done
echo "== after disc tests"
timeout 300 python -m pytest tests/test_disc_gpu.py tests/test_trainer_gen_gpu.py -m gpu -q -p no:cacheprovider 2>&1 | grep -E "^E  |passed|failed" | head -5
