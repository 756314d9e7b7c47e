"""GPU tuning sweep: hot-kernel time vs (rows per tile, load batch); fused vs three-kernel solve."""
import ctypes as C
import json
import os
import sys
import time

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
names = {0: ("cg_calc_w", 32), 1: ("cg_calc_ur", 48), 2: ("cg_calc_p", 24), 3: ("cg_calc_pw", 48)}
out = []
for k in (0, 1, 2, 3):
    for rows in (8, 16, 32, 64):
        for batch in (1, 2, 4):
            assert L.tl_set_tuning(k, rows, batch) == 0
            ms = C.c_double()
            rc = L.tl_time_kernel(app.chunk.handle, k, 30, C.byref(ms))
            assert rc == 0, L.tl_last_error()
            gbs = cells * names[k][1] / ms.value / 1e6
            out.append(dict(kernel=names[k][0], rows=rows, batch=batch, ms=ms.value, gb_s=gbs))
            print("%-11s rows=%3d batch=%d  %.4f ms  %7.1f GB/s" % (names[k][0], rows, batch, ms.value, gbs), flush=True)
best = {}
for o in out:
    if o["kernel"] not in best or o["ms"] < best[o["kernel"]]["ms"]:
        best[o["kernel"]] = o
print("BEST", json.dumps(best))
for k, nm in enumerate(("cg_calc_w", "cg_calc_ur", "cg_calc_p", "cg_calc_pw")):
    L.tl_set_tuning(k, best[nm]["rows"], best[nm]["batch"])
app.close()
# full solves: three-kernel vs fused, 600 iterations
for fused in (False, True):
    s, st = read_config(os.path.join(root, "tests", "decks", "tea_4000_cg.in"), Settings(grid_x_cells=n, grid_y_cells=n))
    s.max_iters = 600
    s.fuse_p_into_w = fused
    app = TeaLeaf(s, st)
    app.solve(0)
    info = app.solve(1)
    print("fused=%d iters=%d gpu_ms=%.2f ms/iter=%.4f cell-iter/s=%.4e GB/s(104B)=%.1f temp=%r" % (
        fused, info.total_iters, info.gpu_ms, info.gpu_ms / info.total_iters,
        cells * info.total_iters / info.gpu_ms * 1e3, cells * info.total_iters * 104 / info.gpu_ms / 1e6,
        app.field_summary_driver()["temp"]), flush=True)
    app.close()
