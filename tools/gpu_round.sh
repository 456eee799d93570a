#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"
tail -6 gpurun_out/pytest_gpu.log
timeout 120 python tools/step_runner.py hwr_train --B 8 --steps 20 --graph 2>&1 | tail -1
timeout 120 python tools/step_runner.py hwr_train --B 32 --steps 20 --graph 2>&1 | tail -1
timeout 120 python tools/step_runner.py gen_train --B 16 --steps 20 --graph 2>&1 | tail -1
