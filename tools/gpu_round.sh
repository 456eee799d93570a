#!/bin/bash
mkdir -p gpurun_out
python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "bench exit $?"
tail -c 300 gpurun_out/bench_default.err
python bench.py --impl reference --steps 6 --warmup 1 > gpurun_out/bench_reference.json 2>/dev/null; echo "ref exit $?"
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
