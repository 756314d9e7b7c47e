"""Short CG run for ncu: one 4000x4000 timestep capped at --iters CG iterations (never a bench number)."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from exploringsycl_b200 import Settings, TeaLeaf, read_config  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=4000)
ap.add_argument("--iters", type=int, default=20)
ap.add_argument("--solver", default="cg")
ap.add_argument("--fused", type=int, default=1)
a = ap.parse_args()
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
s, st = read_config(os.path.join(root, "tests", "decks", "tea_4000_cg.in"), Settings(grid_x_cells=a.n, grid_y_cells=a.n))
s.max_iters = a.iters
s.solver = {"jacobi": 0, "cg": 1, "cheby": 2, "ppcg": 3}[a.solver]
s.fuse_p_into_w = a.fused
app = TeaLeaf(s, st)
info = app.solve(0)
print("iters", info.total_iters, "gpu_ms", info.gpu_ms, app.field_summary_driver())
app.close()
