"""dynamic instruction mix of one profiled launch: python scripts/ncu_mix.py rep.ncu-rep [launch_index]"""
import csv, subprocess, sys, collections
rep = sys.argv[1]
skip = sys.argv[2] if len(sys.argv) > 2 else "0"
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass", "--launch-skip", skip,
                      "--launch-count", "1"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
for i, r in enumerate(rows):
    if "Source" in r and any("Sampl" in c for c in r):
        hdr, start = r, i + 1
        break
si = hdr.index("Source")
ei = hdr.index("Instructions Executed")
ki = hdr.index("# Samples")
mix, stall = collections.Counter(), collections.Counter()
tot = 0
for r in rows[start:]:
    try:
        n = int(r[ei]); s = int(r[ki])
    except (ValueError, IndexError):
        continue
    toks = r[si].split()
    op = toks[1] if toks and toks[0].startswith("@") and len(toks) > 1 else (toks[0] if toks else "?")
    op = ".".join(op.split(".")[:2]) if op.startswith(("LD", "ST", "SHFL", "SYNCS")) else op.split(".")[0]
    mix[op] += n; stall[op] += s; tot += n
print("total warp instructions", tot, "stall samples", sum(stall.values()))
for op, n in mix.most_common(40):
    print(f"{op:24s} {n:12d} {100*n/tot:5.1f}%   stall {100*stall[op]/max(1,sum(stall.values())):5.1f}%")
