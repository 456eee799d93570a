#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"
tail -3 gpurun_out/pytest_gpu.log
timeout 100 python tools/stream_bench.py blur maxpool_bwd 2>&1 | tail -4
timeout 120 python tools/step_runner.py gen_train --B 16 --steps 20 --graph 2>&1 | tail -1
timeout 200 python bench.py --workload hwr_train --steps 20 > gpurun_out/bench_hwr_train.json 2> gpurun_out/bench_hwr_train.err; echo "hwr bench exit $?"; tail -c 300 gpurun_out/bench_hwr_train.err; cut -c1-200 gpurun_out/bench_hwr_train.json
