"""per-kernel HBM table from the ncu csv of scripts/kernel_zoo.py:
python scripts/kernel_zoo_table.py zoo.csv [peak_GBs]"""
import collections
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
peak = float(sys.argv[2]) if len(sys.argv) > 2 else None
if peak is None:
    try:
        peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    except (OSError, KeyError):
        peak = 6431.1
rows = [r for r in csv.reader(open(sys.argv[1], errors="replace")) if len(r) > 10]
hdr = rows[0]
ki, mi, ui, vi, ii = (hdr.index(k) for k in ("Kernel Name", "Metric Name", "Metric Unit", "Metric Value", "ID"))
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1.0,
        "usecond": 1e-6, "nsecond": 1e-9, "msecond": 1e-3, "second": 1.0}
launch = collections.OrderedDict()
for r in rows[1:]:
    d = launch.setdefault(r[ii], {"name": r[ki]})
    d[r[mi]] = float(r[vi].replace(",", "")) * UNIT.get(r[ui], 1.0)
agg = collections.OrderedDict()
for d in launch.values():
    if not d["name"].startswith(("nkb::", "void nkb::")):
        continue
    a = agg.setdefault(d["name"], [0, 0.0, 0.0, 0.0, 0.0])
    t = d.get("gpu__time_duration.sum", 0.0)
    rd, wr = d.get("dram__bytes_read.sum", 0.0), d.get("dram__bytes_write.sum", 0.0)
    a[0] += 1
    a[1] += t
    a[2] += rd
    a[3] += wr
    a[4] = max(a[4], (rd + wr) / t / 1e9 if t > 0 else 0.0)
print(f"per-kernel DRAM throughput (ncu dram__bytes_read.sum + dram__bytes_write.sum over gpu__time_duration.sum,\n"
      f"--clock-control none; cold-cache, serialised launches) against the measured HBM peak {peak:.1f} GB/s\n")
print(f"{'kernel':72s} {'n':>3s} {'time ms':>9s} {'read GB':>8s} {'write GB':>8s} {'GB/s':>7s} {'of peak':>7s} {'best':>7s}")
for name, (n, t, rd, wr, best) in agg.items():
    gbs = (rd + wr) / t / 1e9 if t > 0 else 0.0
    print(f"{name[:72]:72s} {n:3d} {t*1e3:9.3f} {rd/1e9:8.3f} {wr/1e9:8.3f} {gbs:7.0f} {gbs/peak:7.2f} {best/peak:7.2f}")
