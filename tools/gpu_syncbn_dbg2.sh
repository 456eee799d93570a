#!/bin/bash
# two-GPU SyncBN: check tool + bench with HWG_BENCH_SYNC_BN=1, non-blocking; on failure rerun with per-launch syncs
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 200 $TR --master-port 29513 tools/dp_syncbn_check.py > gpurun_out/syncbn_w2.log 2>&1; echo "w2 check exit $?"
grep "SYNCBN\|SyncBN\|illegal" gpurun_out/syncbn_w2.log | head -5
if ! grep -q "SYNCBN_CHECK PASS" gpurun_out/syncbn_w2.log; then
  HWG_DEBUG_SYNC=1 timeout 200 $TR --master-port 29514 tools/dp_syncbn_check.py > gpurun_out/syncbn_w2_dbg.log 2>&1; echo "w2 debug-sync exit $?"
  grep -v "^frame\|^\*\|OMP_NUM" gpurun_out/syncbn_w2_dbg.log | grep -v "^$" | head -40
fi
export HWG_BENCH_NO_EXTRAS=1
HWG_BENCH_SYNC_BN=1 timeout 300 $TR --master-port 29511 bench.py --gpus 2 --steps 30 --warmup 3 > gpurun_out/bench_dp2_syncbn.json 2> gpurun_out/bench_dp2_syncbn.err; echo "dp2 syncbn bench exit $?"
python -c "
import json; d=json.loads(open('gpurun_out/bench_dp2_syncbn.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['config']['batchnorm'][:40], d['config']['execution'][:80])"
grep -v "^frame\|^\*\|OMP_NUM" gpurun_out/bench_dp2_syncbn.err | grep -v "^$" | head -20
