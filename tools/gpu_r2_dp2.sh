#!/bin/bash
# Round 2: the data-parallel balanced step on 2 GPUs (strong scaling: 64 lines per GPU): does it run, stay in sync, how fast
mkdir -p gpurun_out
exec > >(tee gpurun_out/r2_dp2.log) 2>&1
nvidia-smi -L
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 \
  bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_r2_dp2.json 2> gpurun_out/bench_r2_dp2.err
tail -c 1500 gpurun_out/bench_r2_dp2.json; tail -15 gpurun_out/bench_r2_dp2.err
echo "== weak variant: 16 lines per GPU, SyncBN through NCCL"
HWG_BENCH_B=16 HWG_BENCH_SYNC_BN=nccl timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 \
  bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_r2_dp2_b16.json 2> gpurun_out/bench_r2_dp2_b16.err
tail -c 600 gpurun_out/bench_r2_dp2_b16.json; tail -5 gpurun_out/bench_r2_dp2_b16.err
echo "== 2-rank tests"
timeout 600 python -m pytest tests/test_peer_gpu.py -m gpu -q -p no:cacheprovider 2>&1 | tail -5
