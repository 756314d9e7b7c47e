"""Rows-per-tile / load-batch sweep of the stencil kernels (calc_w, fused calc_pw, one-pass Chebyshev / PPCG), 4000x4000."""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from exploringsycl_b200 import Settings, TeaLeaf, lib, read_config  # noqa: E402

root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4000
L = lib()
s, st = read_config(os.path.join(root, "tests", "decks", "tea_4000_cg.in"), Settings(grid_x_cells=n, grid_y_cells=n))
s.max_iters = 40
app = TeaLeaf(s, st)
app.solve(0)
cells = n * n


def t(which, reps=30):
    ms = C.c_double()
    assert L.tl_time_kernel(app.chunk.handle, which, reps, C.byref(ms)) == 0, L.tl_last_error()
    return ms.value


for k, which, name, bpc in ((0, 0, "cg_calc_w", 32), (3, 3, "cg_calc_pw", 48)):
    for rows in (0, 16, 28, 40, 48, 55, 64, 87):
        for batch in (1, 2):
            assert L.tl_set_tuning(k, rows, batch) == 0
            ms = t(which)
            print("%-11s rows=%3d batch=%d  %.4f ms  %7.1f GB/s" % (name, rows, batch, ms, cells * bpc / ms / 1e6), flush=True)
    L.tl_set_tuning(k, 0, 2 if k == 0 else 1)
for rows in (0, 40, 55, 64, 87, 100, 128):  # the one-pass kernels take their rows from the calc_w setting when it is > 0
    L.tl_set_tuning(0, rows, 2)
    for which, name in ((16, "cheby fused"), (17, "ppcg fused")):
        ms = t(which)
        print("%-11s rows=%3d          %.4f ms  %7.1f GB/s" % (name, rows, ms, cells * 64 / ms / 1e6), flush=True)
app.close()
