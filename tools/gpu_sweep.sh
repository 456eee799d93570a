#!/bin/bash
for v in 296 148 96 64 32; do
  echo "HWG_WGRAD_CTAS=$v"; HWG_WGRAD_CTAS=$v python tools/step_runner.py gen_train --B 16 --steps 20 --graph 2>&1 | tail -1
done
