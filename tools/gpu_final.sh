#!/bin/bash
# round-end evidence on one GPU (artifacts kept small: gpurun_out/ is capped at 64 MiB): full GPU test suite, the
# default bench line and the reference arm
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"
tail -2 gpurun_out/pytest_gpu.log
timeout 900 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "bench exit $?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_default.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], 'launches/step', d['gpu_launches']/d['steps'], 'traffic', d['roofline']['traffic'])
for r in [d['roofline']]+d['roofline_other_kernels']:
    print(r['kernel'], round(r['achieved'],1), r['unit'], round(r['frac'],3), 'ms', round(r['kernel_ms_per_step'],3), 'n', r['launches_per_step'])
print(d['cpu_baseline']); print(d.get('clocks')); print({k:(round(v['ms_per_step'],3), round(v['lines_per_s'],1)) for k,v in d.get('extra_workloads',{}).items() if isinstance(v,dict)})
PY
