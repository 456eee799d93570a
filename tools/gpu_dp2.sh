#!/bin/bash
mkdir -p gpurun_out
export HWG_BENCH_NO_EXTRAS=1
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 30 --warmup 3 > gpurun_out/bench_dp2.json 2> gpurun_out/bench_dp2.err; echo "dp2 exit $?"
tail -c 1500 gpurun_out/bench_dp2.err
cut -c1-400 gpurun_out/bench_dp2.json
