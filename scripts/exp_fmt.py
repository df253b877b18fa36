import json
import sys

for line in sys.stdin:
    if line.startswith("{"):
        d = json.loads(line)
        S = d["config"]["time_steps_per_year"]
        print("evals/s(at S=2400) %.1f frac %.4f avg_launch_ms %.4f" % (d["value"] * S / 2400, d["roofline"]["frac"], d["roofline"]["avg_launch_ms"]))
    else:
        print(line.strip()[:300])
