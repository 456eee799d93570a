#!/bin/bash
# multi-GPU evidence (NPROC ranks, default 2): SyncBN parity check (peer-memory exchange and NCCL), then the train-step bench with
# BatchNorm statistics exchanged in-kernel / by NCCL / per rank
mkdir -p gpurun_out
NP=${NPROC:-2}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $NP --master-addr 127.0.0.1"
if [ -z "$SKIP_CHECK" ]; then
  timeout 300 $TR --master-port 29513 tests/tools/dp_syncbn_check.py > gpurun_out/syncbn_w2.log 2>&1; echo "w2 check exit $?"
  grep -v "^frame\|^\*\|OMP_NUM\|^$" gpurun_out/syncbn_w2.log | tail -${CHECK_TAIL:-14}
fi
export HWG_BENCH_NO_EXTRAS=1
for mode in ${MODES:-peer nccl off}; do
  HWG_BENCH_SYNC_BN=$mode timeout 300 $TR --master-port 29511 bench.py --gpus $NP --steps 30 --warmup 3 > gpurun_out/bench_dp2_$mode.json 2> gpurun_out/bench_dp2_$mode.err; echo "dp2 $mode bench exit $?"
  python -c "
import json; d=json.loads(open('gpurun_out/bench_dp2_$mode.json').read().strip().splitlines()[-1]); print('$mode', round(d['value'],1), round(d['ms_per_step'],4), d['gpu_launches'], d['config']['batchnorm'][:60])" || grep -v "^frame\|^\*\|OMP_NUM\|^$" gpurun_out/bench_dp2_$mode.err | tail -15
done
