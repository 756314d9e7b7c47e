"""Resident CG loop rate at n x n (capped iterations) for a list of p+w kernel settings: MODE:STAGES:CTAS:ROWS:FUSED ..."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from exploringsycl_b200 import Settings, TeaLeaf, lib, read_config  # noqa: E402

root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
L = lib()
n = int(sys.argv[1])
for spec in sys.argv[2:]:
    mode, stages, ctas, rows, fused = (int(v) for v in spec.split(":"))
    L.tl_set_pw_pipeline(mode, stages, ctas)
    L.tl_set_tuning(0, rows, 2)
    L.tl_set_tuning(3, rows, 1)
    s, st = read_config(os.path.join(root, "tests", "decks", "tea_4000_cg.in"), Settings(grid_x_cells=n, grid_y_cells=n))
    s.max_iters = 600
    s.fuse_p_into_w = fused
    app = TeaLeaf(s, st)
    best = None
    for t in range(3):
        info = app.solve(t)
        ms = info.gpu_ms / info.total_iters
        best = ms if best is None or ms < best else best
    print("PDL=%s mode=%d stages=%d ctas=%d rows=%2d fused=%d  loop %.4f ms/iter  %.4e cell-iter/s  temp %r" % (
        os.environ.get("TL_PDL", "1"), mode, stages, ctas, rows, fused, best, n * n / best * 1e3,
        app.field_summary_driver()["temp"]), flush=True)
    app.close()
