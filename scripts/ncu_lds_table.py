#!/usr/bin/env python
"""per-instruction shared-memory wavefronts from an `ncu --page source --csv` export: groups the LDS / STS / SHFL
instructions of the hot kernel by opcode and wavefronts per execution (actual vs ideal)"""
import csv
import sys
from collections import defaultdict

rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
agg = defaultdict(lambda: [0, 0, 0, 0])
for r in rows[2:]:
    if len(r) < len(hdr):
        continue
    src = r[ix["Source"]]
    try:
        ex = float(r[ix["Instructions Executed"]] or 0)
        wf = float(r[ix["L1 Wavefronts Shared"]] or 0)
        wi = float(r[ix["L1 Wavefronts Shared Ideal"]] or 0)
    except ValueError:
        continue
    if wf == 0 or ex == 0:
        continue
    op = src.split()[0] if not src.startswith("@") else src.split()[1]
    key = (op, round(wf / ex, 2), round(wi / ex, 2))
    a = agg[key]
    a[0] += 1
    a[1] += ex
    a[2] += wf
    a[3] += wi
tot = sum(a[2] for a in agg.values())
print(f"total shared wavefronts {tot:.3e}")
for key, a in sorted(agg.items(), key=lambda kv: -kv[1][2]):
    print(f"{key[0]:28s} wf/exec {key[1]:5.2f} ideal {key[2]:5.2f}  n_instr {a[0]:4d}  execs {a[1]:.3e}  wavefronts {a[2]:.3e} ({100 * a[2] / tot:.1f} %)")
