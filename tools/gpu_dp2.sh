#!/bin/bash
mkdir -p gpurun_out
export HWG_BENCH_NO_EXTRAS=1
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 tools/dp_syncbn_check.py 2>&1 | grep -v "^\*\|OMP_NUM" | tail -4
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 30 --warmup 3 > gpurun_out/bench_dp2.json 2> gpurun_out/bench_dp2.err; echo "dp2 syncbn exit $?"
python -c "
import json; d=json.loads(open('gpurun_out/bench_dp2.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['config']['batchnorm'][:40], d['config']['execution'][:60])"
HWG_BENCH_NO_SYNC_BN=1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 30 --warmup 3 > gpurun_out/bench_dp2_nosync.json 2> gpurun_out/bench_dp2_nosync.err; echo "dp2 nosync exit $?"
python -c "
import json; d=json.loads(open('gpurun_out/bench_dp2_nosync.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['config']['batchnorm'][:40])"
tail -c 600 gpurun_out/bench_dp2.err
