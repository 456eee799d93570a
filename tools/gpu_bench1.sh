#!/bin/bash
# one-GPU: discriminator tests, then the quick train-step bench with the discriminator branch on a parallel stream
# and serialised
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_disc_gpu.py -q 2>&1 | tail -5
export HWG_BENCH_NO_EXTRAS=1
for tag in par ser; do
  [ $tag = ser ] && export HWG_BENCH_NO_OVERLAP=1
  timeout 400 python bench.py --steps 40 --warmup 3 > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err; echo "bench $tag exit $?"
  tail -c 600 gpurun_out/bench_$tag.err
  python - <<PY
import json
d=json.loads(open('gpurun_out/bench_$tag.json').read().strip().splitlines()[-1])
print('$tag', round(d['value'],1), round(d['ms_per_step'],4), 'e2e', round(d['e2e']['value'],1), d['config']['execution'][:110], 'loss', d['final_loss'])
r=d['roofline']; print(r['kernel'], round(r['achieved'],1), round(r['frac'],3), round(r['kernel_ms_per_step'],3))
PY
done
