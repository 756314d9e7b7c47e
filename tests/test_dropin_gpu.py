"""Drop-in proof: the UNMODIFIED reference host (main.c / diffuse.c / drivers/*.c compiled from
/root/reference by c_kernels/cuda/build_dropin.py into oracle/_ref/) drives this backend through the
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
BIN = {k: os.path.join(ROOT, "oracle", "_ref", k) for k in ("tealeaf_cuda", "tealeaf_cuda_resident")}


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
