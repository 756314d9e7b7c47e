"""python -m exploringsycl_b200: the application flow with the command-line / deck hygiene of SURVEY.md 8f-4 and the
reporting of 8f-3, on the GPU.  -x / -y must really resize the mesh (the reference's main.c:76-85 cannot), the run must
pass the reference's own tea.problems check for that mesh, and --visit must dump the interior fields."""
import json
import os
import re
import shutil
import subprocess
import sys

import numpy as np
import pytest

from oracle import oracle as O
from tl_testutil import DECKS, GOLDEN, ROOT, rel

pytestmark = pytest.mark.gpu


def run_cli(tmp_path, *args):
    env = dict(os.environ, PYTHONPATH=ROOT)
    for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK"):
        env.pop(k, None)
    shutil.copy(os.path.join(GOLDEN, "tea_problems.txt"), tmp_path / "tea.problems")
    r = subprocess.run([sys.executable, "-m", "exploringsycl_b200"] + list(args), cwd=tmp_path, capture_output=True,
                       text=True, timeout=600, env=env)
    return r


def test_cli_resizes_with_x_y_and_reports(tmp_path):
    # the 250x250, 10-step deck turned into the 64x64, 1-step problem of tea.problems row "64 64 1" from the command line
    deck = open(os.path.join(DECKS, "tea_250_cg.in")).read().replace("end_step=10", "end_step=1")
    (tmp_path / "tea.in").write_text(deck)
    r = run_cli(tmp_path, "-x", "64", "-y", "64", "--visit")
    assert r.returncode == 0, r.stdout + r.stderr
    assert "x_cells = 64" in r.stdout and "PASSED" in r.stdout
    ores = O.run_deck(O.make_deck(64, end_step=1))
    cg = [int(v) for v in re.findall(r"(?m)^CG:\s+(\d+) iterations", r.stdout)]
    assert len(cg) == 1 and abs(cg[0] - ores["iters_a"][0]) <= 1
    side = json.load(open(tmp_path / "tea.json"))
    assert side["grid"] == [64, 64] and len(side["steps"]) == 1 and side["steps"][0]["cell_iters_per_s"] > 0
    for key, name in (("vol", "volume"), ("mass", "mass"), ("ie", "internal_energy"), ("temp", "temperature")):
        assert rel(side["field_summary"][name], ores[key]) < 1e-10
    # VisIt bricks: header + 64 x 64 doubles; temperature = u of the oracle run
    u = np.fromfile(tmp_path / "temperature1.dat", dtype="<f8").reshape(64, 64)
    ofields = O.run_deck(O.make_deck(64, end_step=1), want_fields=True)
    assert np.allclose(u, ofields["u"], rtol=1e-9, atol=0.0)
    assert "DATA_SIZE: 64 64 1" in open(tmp_path / "temperature1.bov").read()
    assert os.path.exists(tmp_path / "density1.bov") and os.path.exists(tmp_path / "energy1.dat")


def test_cli_reference_quirks_mode_keeps_the_deck_mesh(tmp_path):
    """--reference-quirks: as in main.c:25-29 the options are applied after the mesh is built; -x cannot resize it and
    leaves 0 in grid_x_cells, so the tea.problems lookup fails exactly as it does for the reference host."""
    shutil.copy(os.path.join(DECKS, "tea_10_cg.in"), tmp_path / "tea.in")
    r = run_cli(tmp_path, "--reference-quirks", "-x", "64", "-s", "jacobi")
    assert r.returncode == 0, r.stdout + r.stderr
    assert "x_cells = 10" in r.stdout and "Jacobi:" in r.stdout
    assert "Problem was not found" in r.stdout


def test_cli_rejects_bad_input(tmp_path):
    (tmp_path / "tea.in").write_text("*tea\nstate 1 density=1.0 energy=1.0\nx_cells=-4\n*endtea\n")
    r = run_cli(tmp_path)
    assert r.returncode == 2 and "positive" in r.stdout
    r = run_cli(tmp_path, "--deck", "missing.in")
    assert r.returncode == 2
