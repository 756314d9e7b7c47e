"""In-loop sweep of the tall-tile height of the fused CG iteration (calc_pw + calc_ur) -- the number that matters is the
resident loop's time per iteration, not the isolated kernel's.  1 rank or N ranks (torchrun; weak-scaling mesh)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from exploringsycl_b200 import Comms, Settings, TeaLeaf, lib, read_config  # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WEAK = {1: (4000, 4000), 2: (4000, 8000), 4: (8000, 8000), 8: (8000, 16000)}
rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
local = int(os.environ.get("LOCAL_RANK", "0"))
L = lib()
nx, ny = WEAK[world]
argv = sys.argv[1:]
if argv and argv[0] == "--cells":  # --cells X Y : another mesh than the weak-scaling one
    nx, ny = int(argv[1]), int(argv[2])
    argv = argv[3:]
specs = argv or ["0:1"]
for n_, spec in enumerate(specs):
    rows, batch = (int(v) for v in spec.split(":"))
    L.tl_set_tuning(3, rows, batch)
    comms = Comms("tune_%s_%d_%d" % (os.environ.get("MASTER_PORT", "0"), os.getppid(), n_), rank, world, device=local) \
        if world > 1 else None
    s, st = read_config(os.path.join(ROOT, "tests", "decks", "tea_4000_cg.in"), Settings(grid_x_cells=nx, grid_y_cells=ny))
    s.max_iters, s.fuse_p_into_w = 600, 2
    app = TeaLeaf(s, st, comms, device=local)
    best = None
    for t in range(3):
        info = app.solve(t)
        ms = info.gpu_ms / info.total_iters
        best = ms if best is None or ms < best else best
    if rank == 0:
        print("ranks=%d mesh %dx%d calc_pw rows=%3d batch=%d  loop %.4f ms/iter  %.4e cell-iter/s" % (
            world, nx, ny, rows, batch, best, nx * ny / best * 1e3), flush=True)
    app.close()
    if comms:
        comms.finalise()
