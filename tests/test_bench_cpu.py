"""bench.py pieces that run without a GPU: the reference arm (CPU oracle port) prints one JSON line with the
contract's keys, and non-zero ranks of a torchrun launch exit without work."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run(args, env_extra=None):
    env = dict(os.environ)
    env["OMP_NUM_THREADS"] = "2"
    env.update(env_extra or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + args, capture_output=True, text=True,
                          timeout=300, env=env, cwd=ROOT)


def test_reference_arm_json_line():
    r = run(["--impl", "reference", "--steps", "2", "--warmup", "1", "--mesh", "64", "64"])
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "CG cell-iterations/s" and d["unit"] == "cell-iter/s"
    assert d["higher_is_better"] is True and d["dtype"] == "f64" and d["data"] == "synthetic"
    assert d["value"] > 0 and d["steps"] == 2 and d["warmup"] == 1 and d["vs_baseline"] is None
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": "cell-iter/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"]


def test_reference_arm_other_ranks_do_nothing():
    r = run(["--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0"], {"RANK": "1", "WORLD_SIZE": "2"})
    assert r.returncode == 0 and r.stdout.strip() == ""
