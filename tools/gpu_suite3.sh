#!/bin/bash
# the whole GPU suite several times in a row (flakiness check: statistics are summed with fp32 atomics)
mkdir -p gpurun_out
exec > >(tee gpurun_out/suite3.log) 2>&1
for i in $(seq 1 ${RUNS:-3}); do
  timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider 2>&1 | grep -E "^E  .*(Assertion|assert)|^FAILED|passed|failed" | head -12
done
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
