"""bench.py's reference arm on the CPU (no GPU needed): the JSON line of the contract on a small workload, and the
bounded wait that keeps a workload whose Radau steps take minutes from holding the run"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(extra_env, *args):
    env = dict(os.environ, **extra_env)
    for key in ("RANK", "WORLD_SIZE", "LOCAL_RANK"):
        env.pop(key, None)
    res = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", *args], env=env,
                         capture_output=True, text=True, timeout=600, check=False)
    assert res.returncode == 0, res.stderr[-2000:]
    lines = [ln for ln in res.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1, res.stdout[-2000:]
    return json.loads(lines[0]), res.stderr


def test_reference_arm_line_on_a_small_workload():
    line, _ = _run({}, "--grid", "ci30x30", "--module", "iage", "--steps", "1", "--warmup", "0")
    assert line["impl"] == "reference" and line["higher_is_better"] is True and line["vs_baseline"] is None
    assert line["metric"] == "model-year evals/sec (batched perturbations)" and line["unit"] == "model-year evals/s"
    assert line["value"] > 0 and line["config"]["grid"] == "ci30x30" and line["config"]["module"] == "iage"
    cpu = line["cpu_baseline"]
    assert cpu["kind"] == "port" and cpu["cores"] >= 1 and cpu["value"] == line["value"] and cpu["sample"]
    assert line["e2e"] == {"value": line["value"], "unit": line["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_reports_a_workload_it_cannot_sample_in_time():
    line, err = _run({"NKB_CPU_LEG_TIMEOUT_S": "1"}, "--grid", "default40x50", "--module", "iage", "--steps", "1",
                     "--warmup", "0")
    assert line["impl"] == "reference" and "did not finish within 1 s" in line["unavailable"]
    assert "CPU reference leg unavailable" in err
