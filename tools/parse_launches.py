"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list (development aid)."""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
hdr = [i for i, r in enumerate(rows) if r and r[0] == 'ID'][0]
h = rows[hdr]
ki, vi, gi = h.index('Kernel Name'), h.index('Metric Value'), h.index('Grid Size')
first = int(sys.argv[2]) if len(sys.argv) > 2 else 0
last = int(sys.argv[3]) if len(sys.argv) > 3 else 10**9
agg = collections.OrderedDict()
seq = []
for r in rows[hdr + 1:]:
    if len(r) <= vi: continue
    i = int(r[0])
    if i < first or i >= last: continue
    name = r[ki].split('(')[0][-48:]
    t = float(r[vi].replace(',', '')) / 1e3
    seq.append((i, name, r[gi], t))
    a = agg.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += t
tot = sum(a[1] for a in agg.values())
for n, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{n:50s} n={c:4d} total={t:9.1f} us  {100 * t / tot:5.1f}%")
print(f"TOTAL {tot:.1f} us over {len(seq)} launches")
if '-v' in sys.argv:
    for i, n, g, t in seq: print(i, n, g, f"{t:.1f}")
