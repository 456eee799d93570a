"""Copies the outputs of tools/gpu_r2_final.sh from gpurun_out/ into profiles/ and prints the numbers DESIGN.md quotes."""
import json, shutil, subprocess
def last(p):
    L = [l for l in open(p).read().splitlines() if l.startswith('{')]
    return json.loads(L[-1]), L[-1]
b, l = last('gpurun_out/bench_r02_final_b128.json')
open('profiles/bench_r02_final_b128.json', 'w').write(l + "\n")
print('B128', round(b['value'], 1), round(b['ms_per_step'], 2), 'e2e', round(b['e2e']['value'], 1), 'eager', round(b['eager_ms_per_step'], 2), b['clocks'], 'launches', b['gpu_launches'])
r = b['roofline']
print('roof', r['kernel'], round(r['achieved'], 1), round(r['frac'], 3), round(r['kernel_ms_per_step'], 2), r['traffic'], r['launches_per_step'], round(r['algorithmic_gflop_per_step'], 1))
for o in b['roofline_other_kernels']:
    print(' ', o['kernel'], round(o['achieved'], 1), round(o['frac'], 3), round(o['kernel_ms_per_step'], 2), o['traffic'])
print('gpu_baseline', round(b['gpu_baseline']['value'], 1), round(b['gpu_baseline']['ms_per_step'], 1), 'cpu', b['cpu_baseline']['value'])
for k, v in b['extra_workloads'].items():
    print(' ', k, {kk: round(vv, 2) for kk, vv in v.items() if kk in ('ms_per_step', 'lines_per_s', 'ms_per_cycle')})
b2, l2 = last('gpurun_out/bench_r02_final_b16.json'); open('profiles/bench_r02_final_b16.json', 'w').write(l2 + "\n")
print('B16', round(b2['value'], 1), round(b2['ms_per_step'], 3), round(b2['roofline']['frac'], 3), round(b2['eager_ms_per_step'], 2))
b3, l3 = last('gpurun_out/bench_r02_final_reference.json'); open('profiles/bench_r02_final_reference.json', 'w').write(l3 + "\n")
print('ref', b3.get('impl'), round(b3['value'], 2), round(b3['ms_per_step'], 1))
shutil.copy('gpurun_out/launches_gan_step_b128_final.txt', 'profiles/launches_gan_step_b128_r02_final.txt')
shutil.copy('gpurun_out/launches_gan_step_b128_final.csv.gz', 'profiles/launches_gan_step_b128_r02_final.csv.gz')
