"""Tiny 2-rank repro of the resident multi-rank CG loop (debug aid; run under `timeout`)."""
import os
import sys
import torch.multiprocessing as mp

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def worker(rank, world, session, n, iters):
    from exploringsycl_b200 import Comms, Settings, TeaLeaf, read_config
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    comms = Comms(session, rank, world, device=rank)
    s, st = read_config(os.path.join(root, "tests", "decks", "tea_250_cg.in"), Settings(grid_x_cells=n, grid_y_cells=n))
    s.max_iters = iters
    s.batch = 4
    app = TeaLeaf(s, st, comms, device=rank)
    print("rank", rank, "decomp", app.decomposition, flush=True)
    try:
        info = app.solve(0)
        print("rank", rank, "iters", info.total_iters, "err", info.error, app.field_summary_driver(), flush=True)
    except Exception as e:
        print("rank", rank, "FAILED", e, flush=True)
    comms.barrier()
    app.close()
    comms.finalise()


if __name__ == "__main__":
    world = int(sys.argv[1]) if len(sys.argv) > 1 else 2
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 64
    iters = int(sys.argv[3]) if len(sys.argv) > 3 else 6
    ctx = mp.get_context("spawn")
    ps = [ctx.Process(target=worker, args=(r, world, "dbg%d" % os.getpid(), n, iters)) for r in range(world)]
    [p.start() for p in ps]
    [p.join() for p in ps]
