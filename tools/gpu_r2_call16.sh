#!/bin/bash
# full GPU suite + smoke on the current library
mkdir -p gpurun_out
exec > >(tee gpurun_out/r2_call16.log) 2>&1
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider 2>&1 | grep -v "^$" | tail -15
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3
