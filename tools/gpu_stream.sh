#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python tools/stream_bench.py 2>&1 | tail -12
