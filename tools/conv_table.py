"""Per-layer table from bench.py's HWG_BENCH_DUMP_CONV file (development aid): launches grouped by (kernel, geometry)."""
import collections
import json
import sys

d = json.load(open(sys.argv[1]))
agg = collections.OrderedDict()
for r in d["launches"]:
    k = (r["kind"], tuple(r["geom"] or ()))
    a = agg.setdefault(k, [0, 0.0, 0.0, 0.0])
    a[0] += 1; a[1] += r["ms"]; a[2] += r["gflop"]; a[3] += r["mb"]
steps = d["steps"]
tot = sum(a[1] for a in agg.values()) / steps
print(f"batch {d['batch']}: {tot:.2f} ms of convolution launches per step")
for (kind, geom), (n, ms, gf, mb) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{kind:18s} N,Ho,Wo,Cin,Cout,taps={str(geom):34s} x{n / steps:4.1f}  {ms / steps * 1e3:8.1f} us/step  "
          f"{gf / ms:7.1f} TFLOP/s  {mb / ms:7.1f} GB/s")
