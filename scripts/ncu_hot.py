"""top stalled SASS instructions of one profiled launch: python scripts/ncu_hot.py rep.ncu-rep [launch_index]"""
import csv
import subprocess
import sys

rep = sys.argv[1]
skip = sys.argv[2] if len(sys.argv) > 2 else "0"
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass", "--launch-skip", skip,
                      "--launch-count", "1"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, start = None, 0
for i, r in enumerate(rows):
    if "Source" in r and any("Sampl" in c for c in r):
        hdr, start = r, i + 1
        break
si = hdr.index("Source")
k = [j for j, c in enumerate(hdr) if "Samples" in c][0]
data = []
for r in rows[start:]:
    try:
        data.append((int(r[k]), r[si]))
    except (ValueError, IndexError):
        pass
tot = sum(d[0] for d in data)
print(hdr[k], "total", tot, "instructions", len(data))
idx = sorted(range(len(data)), key=lambda i: -data[i][0])[:int(sys.argv[3]) if len(sys.argv) > 3 else 30]
for i in sorted(idx):
    print(i, data[i][0], "%.1f%%" % (100 * data[i][0] / tot), data[i][1][:120])
