"""Generates tests/golden/ref_host_goldens.json by running the UNMODIFIED reference host
(oracle/_ref/tealeaf_ref = main.c / diffuse.c / drivers/*.c compiled in place from /root/reference, kernels =
the oracle) on the decks of tests/decks/.  Run in the dev container (needs /root/reference at build time):

    python tests/golden/make_ref_host_goldens.py
"""
import json
import os
import re
import shutil
import subprocess
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
REF = os.path.join(ROOT, "oracle", "_ref", "tealeaf_ref")
DECKS = ["tea_10_cg.in", "tea_64_cg.in", "tea_250_cg.in", "tea_500_cg.in", "tea_250_cheby.in", "tea_250_ppcg.in",
         "tea_10_jacobi.in"]

out = {}
for deck in DECKS:
    d = tempfile.mkdtemp()
    shutil.copy(os.path.join(ROOT, "tests", "decks", deck), os.path.join(d, "tea.in"))
    shutil.copy(os.path.join(ROOT, "tests", "golden", "tea_problems.txt"), os.path.join(d, "tea.problems"))
    env = dict(os.environ, OMP_NUM_THREADS="8")
    txt = subprocess.run([REF], cwd=d, capture_output=True, text=True, env=env, timeout=3600).stdout
    out[deck] = {
        "cg": [int(v) for v in re.findall(r"(?m)^CG:\s+(\d+) iterations", txt)],
        "cheby": [int(v) for v in re.findall(r"(?m)^Cheby:\s+(\d+) iterations", txt)],
        "ppcg": [int(v) for v in re.findall(r"(?m)^PPCG:\s+(\d+) iterations", txt)],
        "jacobi": [int(v) for v in re.findall(r"(?m)^Jacobi:\s+(\d+) iterations", txt)],
        "actual_temp": float(re.search(r"Actual\s+(\S+)", txt).group(1)),
        "passed": "PASSED" in txt,
    }
    shutil.rmtree(d)
json.dump(out, open(os.path.join(ROOT, "tests", "golden", "ref_host_goldens.json"), "w"), indent=1, sort_keys=True)
print(json.dumps(out, indent=1)[:600])
