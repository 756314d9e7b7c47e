"""The UNMODIFIED reference host (main.c / diffuse.c / drivers/*.c compiled in place from
/root/reference into oracle/_ref/tealeaf_ref, kernels = the oracle) must print the same iteration
counts and final temperature as the oracle's own restatement of the drivers.  Runs only where the
reference binary exists (it is built in the dev container and travels with the repo snapshot)."""
import os
import re
import shutil
import subprocess

import pytest

from oracle import oracle as O
from tl_testutil import DECKS, GOLDEN, rel

pytestmark = pytest.mark.skipif(not os.path.exists(O.REF_BIN), reason="oracle/_ref/tealeaf_ref not built")


def run_ref(deck, tmp_path, solver_flag=None):
    shutil.copy(os.path.join(DECKS, deck), tmp_path / "tea.in")
    shutil.copy(os.path.join(GOLDEN, "tea_problems.txt"), tmp_path / "tea.problems")
    cmd = [O.REF_BIN] + (["-solver", solver_flag] if solver_flag else [])
    out = subprocess.run(cmd, cwd=tmp_path, capture_output=True, text=True, timeout=600).stdout
    return out


@pytest.mark.parametrize("n", [10, 250])
def test_cg(n, tmp_path):
    out = run_ref("tea_%d_cg.in" % n, tmp_path)
    iters = [int(v) for v in re.findall(r"(?m)^CG:\s+(\d+) iterations", out)]
    actual = float(re.search(r"Actual\s+(\S+)", out).group(1))
    assert "PASSED" in out
    r = O.run_deck(O.make_deck(n))
    assert iters == r["iters_a"]
    assert rel(actual, r["temp"]) < 1e-14


@pytest.mark.parametrize("name,solver", [("cheby", O.CHEBY), ("ppcg", O.PPCG)])
def test_cheby_ppcg(name, solver, tmp_path):
    out = run_ref("tea_250_%s.in" % name, tmp_path)
    label = "Cheby" if solver == O.CHEBY else "PPCG"
    cg = [int(v) for v in re.findall(r"(?m)^CG:\s+(\d+) iterations", out)]
    it = [int(v) for v in re.findall(r"(?m)^" + label + r":\s+(\d+) iterations", out)]
    actual = float(re.search(r"Actual\s+(\S+)", out).group(1))
    r = O.run_deck(O.make_deck(250, solver=solver))
    assert cg == r["iters_a"] and it == r["iters_b"]
    assert rel(actual, r["temp"]) < 1e-14


def test_jacobi(tmp_path):
    out = run_ref("tea_10_jacobi.in", tmp_path)
    it = [int(v) for v in re.findall(r"Jacobi:\s+(\d+) iterations", out)]
    r = O.run_deck(O.make_deck(10, solver=O.JACOBI))
    assert it == r["iters_a"]
