"""bench.py keeps the driver's contract: one JSON line with the metric of BASELINE.json, `roofline`, `e2e`,
`cpu_baseline`, `clocks`, `gpu_launches`; the reference arm reports the same `config`, `metric`, `unit`,
`higher_is_better`.  Run on a small grid so that the test takes seconds."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SMALL = ["--grid", "ci30x30", "--module", "iage", "--members", "64", "--nsteps", "48", "--steps", "2", "--warmup", "1"]


def _line(extra):
    res = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + SMALL + extra, capture_output=True, text=True,
                         timeout=900, cwd=ROOT)
    assert res.returncode == 0, res.stderr[-3000:]
    lines = [ln for ln in res.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1, res.stdout[-2000:]
    return json.loads(lines[0])


def test_bench_json_line_and_reference_arm():
    ours = _line([])
    base = json.load(open(os.path.join(ROOT, "BASELINE.json")))
    assert ours["metric"].split(" at ")[0].split(" (")[0] in base["metric"]
    for key in ("value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "roofline", "e2e", "gpu_launches", "clocks", "cpu_baseline"):
        assert key in ours, key
    assert ours["dtype"] == "f64" and ours["higher_is_better"] is True and ours["vs_baseline"] is None
    assert ours["n_gpus"] == 1 and ours["steps"] == 2 and ours["warmup"] == 1 and ours["value"] > 0
    assert "workload" in ours["config"] and "model" not in ours["config"]
    roof = ours["roofline"]
    assert roof["bound"] == "hbm" and roof["unit"] == "GB/s" and roof["peak"] > 0
    assert abs(roof["frac"] - roof["achieved"] / roof["peak"]) < 1e-9
    e2e = ours["e2e"]
    assert e2e["value"] > 0 and e2e["h2d_bytes_per_step"] == e2e["d2h_bytes_per_step"] == 8 * ours["config"]["N"] * 64
    assert e2e["value"] != ours["value"]  # measured through the host entry point, not a copy of the device number
    assert ours["gpu_launches"] > 0
    assert set(ours["clocks"]) >= {"sm_mhz", "sm_max_mhz", "reasons"}
    cpu = ours["cpu_baseline"]
    assert cpu["kind"] == "port" and cpu["cores"] == 1 and cpu["value"] > 0 and cpu["sample"]
    ref = _line(["--impl", "reference"])
    assert ref["impl"] == "reference" and ref["value"] > 0
    for key in ("metric", "unit", "higher_is_better", "dtype"):
        assert ref[key] == ours[key], key
    assert ref["config"] == ours["config"]
    assert ref["e2e"] == {"value": ref["value"], "unit": ref["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert ref["cpu_baseline"]["value"] == ref["value"] and ref["cpu_baseline"]["cores"] >= 1
    # the GPU path beats the reference's CPU path by orders of magnitude on the same workload
    assert ours["e2e"]["value"] > 100 * ref["value"]
