#!/bin/bash
# bench.py on N GPUs of one box the way the driver launches it (strong scaling at global batch 128); extras / baselines off
N=${NPROC:-2}
mkdir -p gpurun_out
exec > >(tee gpurun_out/r2_dp$N.log) 2>&1
nvidia-smi -L | head -8
HWG_BENCH_NO_EXTRAS=1 HWG_BENCH_NO_CPU_BASELINE=1 HWG_BENCH_NO_GPU_BASELINE=1 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N \
  --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/bench_r2_dp$N.json 2> gpurun_out/bench_r2_dp$N.err
tail -c 1500 gpurun_out/bench_r2_dp$N.json; tail -5 gpurun_out/bench_r2_dp$N.err
