#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"
tail -15 gpurun_out/pytest_gpu.log
HWG_BENCH_NO_EXTRAS=1 python bench.py --steps 30 > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "bench exit $?"
tail -c 400 gpurun_out/bench_default.err
cut -c1-250 gpurun_out/bench_default.json
