"""kernel share of a profiled command from its ncu launch list:
python scripts/launch_share.py launches.csv "command line" > profiles/r01_launch_share_X.txt"""
import collections
import csv
import sys

rows = [r for r in csv.reader(open(sys.argv[1], errors="replace")) if len(r) > 10]
hdr = rows[0]
ki, mi, ui, vi = (hdr.index(k) for k in ("Kernel Name", "Metric Name", "Metric Unit", "Metric Value"))
UNIT = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3, "nsecond": 1e-6, "usecond": 1e-3, "msecond": 1.0, "second": 1e3}
agg = collections.OrderedDict()
for r in rows[1:]:
    if r[mi] != "gpu__time_duration.sum":
        continue
    a = agg.setdefault(r[ki], [0, 0.0])
    a[0] += 1
    a[1] += float(r[vi].replace(",", "")) * UNIT.get(r[ui], 1e-6)
tot = sum(a[1] for a in agg.values())
print("kernel share of the profiled launches (ncu --metrics gpu__time_duration.sum --clock-control none, "
      "cold-cache, serialised):")
print("command:", sys.argv[2] if len(sys.argv) > 2 else "?")
for name, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{t:10.3f} ms  {100 * t / tot:5.1f}%  n={n:4d}  avg={1e3 * t / n:9.1f} us  {name[:100]}")
