"""Drop-in proof: the UNMODIFIED reference host (main.c / diffuse.c / drivers/*.c compiled from
/root/reference by c_kernels/cuda/build_dropin.py into c_kernels/cuda/bin/) drives this backend through the
reference's own kernel_interface.h symbols.  Its printed iteration counts and final temperature must
match the oracle.  Skipped where the prebuilt binaries are absent."""
import os
import re
import shutil
import subprocess

import pytest

from oracle import oracle as O
from tl_testutil import DECKS, GOLDEN, ROOT, rel

pytestmark = pytest.mark.gpu
BIN = {k: os.path.join(ROOT, "c_kernels", "cuda", "bin", k) for k in ("tealeaf_cuda", "tealeaf_cuda_resident")}


def run_bin(name, deck, tmp_path):
    if not os.path.exists(BIN[name]):
        pytest.skip("%s not built (needs /root/reference at build time)" % name)
    shutil.copy(os.path.join(DECKS, deck), tmp_path / "tea.in")
    shutil.copy(os.path.join(GOLDEN, "tea_problems.txt"), tmp_path / "tea.problems")
    env = dict(os.environ)
    for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK"):
        env.pop(k, None)
    r = subprocess.run([BIN[name]], cwd=tmp_path, capture_output=True, text=True, timeout=900, env=env)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    return r.stdout


def parse(out, label):
    cg = [int(v) for v in re.findall(r"(?m)^CG:\s+(\d+) iterations", out)]
    other = [int(v) for v in re.findall(r"(?m)^" + label + r":\s+(\d+) iterations", out)] if label else []
    actual = float(re.search(r"Actual\s+(\S+)", out).group(1))
    return cg, other, actual


@pytest.mark.parametrize("binary", ["tealeaf_cuda", "tealeaf_cuda_resident"])
@pytest.mark.parametrize("n", [10, 250])
def test_cg_deck(binary, n, tmp_path):
    out = run_bin(binary, "tea_%d_cg.in" % n, tmp_path)
    assert "PASSED" in out
    cg, _, actual = parse(out, None)
    r = O.run_deck(O.make_deck(n))
    assert all(abs(a - b) <= 1 for a, b in zip(cg, r["iters_a"])) and len(cg) == 10, (cg, r["iters_a"])
    assert rel(actual, r["temp"]) < 1e-10


@pytest.mark.parametrize("binary", ["tealeaf_cuda", "tealeaf_cuda_resident"])
@pytest.mark.parametrize("name,label,solver", [("cheby", "Cheby", O.CHEBY), ("ppcg", "PPCG", O.PPCG)])
def test_cheby_ppcg_deck(binary, name, label, solver, tmp_path):
    out = run_bin(binary, "tea_250_%s.in" % name, tmp_path)
    assert "PASSED" in out
    cg, it, actual = parse(out, label)
    r = O.run_deck(O.make_deck(250, solver=solver))
    assert cg == r["iters_a"]
    step = 10 if solver == O.CHEBY else 1
    assert all(abs(a - b) <= step for a, b in zip(it, r["iters_b"])), (it, r["iters_b"])
    assert rel(actual, r["temp"]) < (1e-10 if it == r["iters_b"] else 5e-9)


@pytest.mark.parametrize("binary", ["tealeaf_cuda", "tealeaf_cuda_resident"])
def test_jacobi_deck(binary, tmp_path):
    out = run_bin(binary, "tea_10_jacobi.in", tmp_path)
    it = [int(v) for v in re.findall(r"(?m)^Jacobi:\s+(\d+) iterations", out)]
    actual = float(re.search(r"Actual\s+(\S+)", out).group(1))
    r = O.run_deck(O.make_deck(10, solver=O.JACOBI))
    assert it == r["iters_a"]
    assert rel(actual, r["temp"]) < 1e-10


def test_resident_binary_reports_all_four_sums_and_a_json_sidecar(tmp_path):
    """the reference computes vol / mass / ie / temp (field_summary_driver.c:11-30) and prints only
    temp; the DIFFUSE_OVERLOAD binary adds a 'Field summary:' line, a 'Solver rate:' line per step and tea.json --
    the reference's own lines stay."""
    import json
    out = run_bin("tealeaf_cuda_resident", "tea_250_cg.in", tmp_path)
    assert "PASSED" in out and out.count("Solver rate:") == 10
    r = O.run_deck(O.make_deck(250))
    m = re.findall(r"Field summary:\s+vol (\S+) mass (\S+) ie (\S+) temp (\S+)", out)
    assert len(m) == 2  # step 1 (summary_frequency 10) and the final one
    for got, key in zip(m[-1], ("vol", "mass", "ie", "temp")):
        assert rel(float(got), r[key]) < 1e-10, key
    side = json.load(open(tmp_path / "tea.json"))
    assert side["solver"] == "cg" and side["grid"] == [250, 250] and len(side["steps"]) == 10
    assert [s["iters_a"] for s in side["steps"]] == [int(v) for v in re.findall(r"(?m)^CG:\s+(\d+) iterations", out)]
    assert all(s["cell_iters_per_s"] > 0 and s["algorithmic_gb_per_s"] > 0 for s in side["steps"])
    for key, name in (("vol", "volume"), ("mass", "mass"), ("ie", "internal_energy"), ("temp", "temperature")):
        assert rel(side["field_summary"][name], r[key]) < 1e-10
    # TL_REPORT=0: nothing but the reference's lines
    os.environ["TL_REPORT"] = "0"
    try:
        quiet = run_bin("tealeaf_cuda_resident", "tea_10_cg.in", tmp_path)
    finally:
        del os.environ["TL_REPORT"]
    assert "Field summary:" not in quiet and "Solver rate:" not in quiet and "PASSED" in quiet
