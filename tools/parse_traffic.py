"""Summarise an `ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --csv` log of the
train step (development aid): per kernel, launches / time / DRAM bytes of ONE step (delimited by adam_flat_kernel)."""
import collections
import csv
import json
import sys

rows = list(csv.reader(open(sys.argv[1])))
hi = [i for i, r in enumerate(rows) if r and r[0] == 'ID'][0]
h = rows[hi]
ki, mi, vi, ui = h.index('Kernel Name'), h.index('Metric Name'), h.index('Metric Value'), h.index('Metric Unit')
launch = collections.OrderedDict()
for r in rows[hi + 1:]:
    if len(r) <= vi:
        continue
    d = launch.setdefault(int(r[0]), {'name': r[ki].split('(')[0].replace('void ', '').replace('hwg::', '')[-44:]})
    v = float(r[vi].replace(',', ''))
    u = r[ui]
    if 'byte' in u:
        v *= {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}[u]
    elif u == 'ms':
        v *= 1e3
    elif u == 'ns':
        v /= 1e3
    d[r[mi]] = v
ids = sorted(launch)
ad = [i for i in ids if 'adam_flat' in launch[i]['name']]
step_len = int(sys.argv[2]) if len(sys.argv) > 2 else None
if len(ad) >= 2:
    window = [i for i in ids if ad[0] < i <= ad[1]]
else:   # one boundary inside the capture: head of the next step + tail of the previous one
    head = [i for i in ids if i > ad[0]]
    tail = [i for i in ids if i <= ad[0]]
    window = head + tail[len(tail) - (step_len - len(head)):]
agg = collections.OrderedDict()
for i in window:
    d = launch[i]
    k = agg.setdefault(d['name'], [0, 0.0, 0.0, 0.0])
    k[0] += 1
    k[1] += d.get('gpu__time_duration.sum', 0)
    k[2] += d.get('dram__bytes_read.sum', 0)
    k[3] += d.get('dram__bytes_write.sum', 0)
tot = [sum(v[j] for v in agg.values()) for j in (1, 2, 3)]
out = {}
for n, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{n:46s} n={v[0]:3d} {v[1]:8.1f} us  rd {v[2] / 1e6:8.1f} MB  wr {v[3] / 1e6:8.1f} MB  "
          f"{(v[2] + v[3]) / v[1] / 1e6 if v[1] else 0:6.2f} TB/s")
    out[n] = {"launches": v[0], "us": round(v[1], 1), "dram_read_mb": round(v[2] / 1e6, 2), "dram_write_mb": round(v[3] / 1e6, 2)}
print('TOTAL %d launches, %.1f us, read %.1f MB, write %.1f MB' % (len(window), tot[0], tot[1] / 1e6, tot[2] / 1e6))
if len(sys.argv) > 3:
    meta = dict(a.split("=", 1) for a in sys.argv[4:])         # e.g. batch=128 step=balanced
    if "batch" in meta:
        meta["batch"] = int(meta["batch"])
    json.dump(dict(meta, launches_per_step=len(window), total_us=round(tot[0], 1), kernels=out), open(sys.argv[3], 'w'), indent=1)
