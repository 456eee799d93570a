#!/bin/bash
# one-GPU: discriminator tests, then the quick train-step bench (graph) and eager/graph timings of the step
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_disc_gpu.py -q 2>&1 | tail -15
export HWG_BENCH_NO_EXTRAS=1
timeout 600 python bench.py --steps 30 --warmup 3 > gpurun_out/bench_disc.json 2> gpurun_out/bench_disc.err; echo "bench exit $?"
tail -c 1500 gpurun_out/bench_disc.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_disc.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], 'launches/step', d['gpu_launches']/d['steps'], d['config']['execution'][:70])
for r in [d['roofline']]+d['roofline_other_kernels']:
    print(r['kernel'], round(r['achieved'],1), r['unit'], round(r['frac'],3), 'ms', round(r['kernel_ms_per_step'],3), 'n', r['launches_per_step'])
print(d['cpu_baseline'])
PY
