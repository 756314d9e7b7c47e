"""Times every kernel of the path on the 4000x4000 mesh (CUDA events, tl_time_kernel) and prints the
roofline table: algorithmic bytes per cell / time vs the measured HBM peak."""
import ctypes as C
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from exploringsycl_b200 import Settings, TeaLeaf, lib, read_config  # noqa: E402

root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4000
try:
    peak = json.load(open(os.path.join(root, "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    peak = 6650.0
L = lib()
s, st = read_config(os.path.join(root, "tests", "decks", "tea_4000_cg.in"), Settings(grid_x_cells=n, grid_y_cells=n))
s.max_iters = 30
app = TeaLeaf(s, st)
app.solve(0)
K = [(0, "cg_calc_w", 32), (1, "cg_calc_ur", 48), (2, "cg_calc_p", 24), (4, "cheby_iterate", 64),
     (5, "cheby_calc_u", 24), (15, "cheby_init", 56), (6, "ppcg_calc_ur", 56), (7, "ppcg_calc_sd", 24),
     (8, "jacobi_iterate(+copy)", 56), (9, "calculate_residual", 40), (10, "calculate_2norm", 8),
     (11, "field_summary", 32), (12, "cg_init (one pass; 120 B algorithmic, 64 moved)", 120), (13, "finalise", 24), (14, "copy_u", 16),
     (16, "cheby fused iter+calc_u", 64), (17, "ppcg fused ur+sd", 64), (3, "cg_calc_pw (fused)", 48)]
print("# %dx%d mesh, peak = %.0f GB/s (MEASURED_PEAKS.json)" % (n, n, peak))
print("%-24s %8s %10s %9s %6s" % ("kernel", "B/cell", "ms/launch", "GB/s", "frac"))
for which, name, bpc in K:
    ms = C.c_double()
    rc = L.tl_time_kernel(app.chunk.handle, which, 20, C.byref(ms))
    assert rc == 0, L.tl_last_error()
    gbs = n * n * bpc / ms.value / 1e6
    print("%-24s %8d %10.4f %9.1f %6.3f" % (name, bpc, ms.value, gbs, gbs / peak), flush=True)
app.close()
