"""Sweep rows-per-tile for the stencil kernels at 4000x4000."""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from exploringsycl_b200 import Settings, TeaLeaf, lib, read_config  # noqa: E402

root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4000
L = lib()
s, st = read_config(os.path.join(root, "tests", "decks", "tea_4000_cg.in"), Settings(grid_x_cells=n, grid_y_cells=n))
s.max_iters = 30
app = TeaLeaf(s, st)
app.solve(0)
for which, tune, name, bpc in ((3, 3, "cg_calc_pw", 48), (0, 0, "cg_calc_w", 32), (16, 0, "cheby_fused", 64)):
    for batch in (1, 2):
        for rows in (24, 32, 40, 48, 54, 56, 64, 72, 80, 96, 128, 160):
            assert L.tl_set_tuning(tune, rows, batch) == 0
            ms = C.c_double()
            assert L.tl_time_kernel(app.chunk.handle, which, 20, C.byref(ms)) == 0
            print("%-12s batch=%d rows=%3d  %.4f ms %7.1f GB/s" % (name, batch, rows, ms.value, n * n * bpc / ms.value / 1e6), flush=True)
    L.tl_set_tuning(tune, 0, 2 if tune == 0 else 1)
app.close()
