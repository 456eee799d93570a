#!/bin/bash
# one-GPU: generator training tests, then the quick train-step bench (parallel branches + side-stream wgrad)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gen_train_gpu.py tests/test_graphs_gpu.py tests/test_optim_gpu.py tests/test_fullsize_gpu.py -q 2>&1 | tail -5
export HWG_BENCH_NO_EXTRAS=1
timeout 400 python bench.py --steps 40 --warmup 3 > gpurun_out/bench_par.json 2> gpurun_out/bench_par.err; echo "bench exit $?"
tail -c 600 gpurun_out/bench_par.err
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_par.json').read().strip().splitlines()[-1])
print(round(d['value'],1), round(d['ms_per_step'],4), 'e2e', round(d['e2e']['value'],1), d['config']['execution'][:110], 'loss', d['final_loss'])
r=d['roofline']; print(r['kernel'], round(r['achieved'],1), round(r['frac'],3), round(r['kernel_ms_per_step'],3))
PY
