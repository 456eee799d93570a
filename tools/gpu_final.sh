#!/bin/bash
# round-end evidence on one GPU: full GPU test suite, the default bench line (+ reference arm), the per-kernel traffic
# pass of one steady-state step, and a --set full capture of every conv_fprop launch of one step
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"
tail -3 gpurun_out/pytest_gpu.log
timeout 900 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "bench exit $?"
tail -c 600 gpurun_out/bench_default.err
timeout 300 python bench.py --impl reference --steps 6 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; echo "reference exit $?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_default.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], 'launches/step', d['gpu_launches']/d['steps'])
for r in [d['roofline']]+d['roofline_other_kernels']:
    print(r['kernel'], round(r['achieved'],1), r['unit'], round(r['frac'],3), 'ms', round(r['kernel_ms_per_step'],3), 'n', r['launches_per_step'])
print(d['cpu_baseline']); print(d.get('clocks')); print(json.dumps(d.get('extra_workloads'))[:1500])
r=json.loads(open('gpurun_out/bench_reference.json').read().strip().splitlines()[-1]); print('reference', r['value'], r['cpu_baseline']['sample'])
PY
bash tools/gpu_traffic.sh | tail -3
timeout 600 ncu --set full --clock-control none --import-source on -k "regex:conv_fprop" -s 60 -c 60 -o gpurun_out/prof_conv_fprop_r01_step \
  python tools/step_runner.py gen_train --B 16 --steps 1 --warmup 1 > gpurun_out/ncu_conv.log 2>&1
tail -1 gpurun_out/ncu_conv.log; ls -la gpurun_out/*.ncu-rep
