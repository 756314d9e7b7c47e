"""Where does a resident CG iteration spend its time?  %globaltimer stamps of every loop kernel (tl_stamps_enable).

    python tools/stamps.py [--iters 300] [--cells 4000 4000] [--fused 0|1|2] [--tag NAME]
    python -m torch.distributed.run --nproc-per-node N ... tools/stamps.py   (N ranks: weak-scaling mesh of bench.py)

Per kernel of the loop (0 matvec = calc_w / calc_pw, 1 calc_ur, 2 calc_p) four stamps are taken on the device:
  t0 first CTA past its dependency wait     t1 tail CTA holds this rank's partial (all tiles done, grid reduction done)
  t2 all ranks' partials gathered           t3 halo hand-shake with the neighbours done
Reported per rank, median over the iterations (first 5 dropped), in microseconds:
  body   = t1 - t0   the kernel's streaming phase incl. the deterministic grid reduction
  gather = t2 - t1   NVLink all-gather of the partials: ~one-way latency on the LAST rank to arrive, latency + skew on
                     the others (every clock is per GPU, so only same-rank differences are used)
  halo   = t3 - t1   halo hand-shake (kernels that send a halo)
  gap    = next kernel's t0 - this kernel's last stamp: launch / dependency-resolution latency between kernels
and the whole iteration (t0 of the matvec to t0 of the next matvec).  Every rank writes gpurun_out/stamps_<tag>_r<rank>.json;
after a barrier rank 0 merges them into gpurun_out/stamps_<tag>.json and prints the table.
"""
import argparse
import ctypes as C
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from exploringsycl_b200 import Comms, Settings, TeaLeaf, lib, read_config  # noqa: E402
from exploringsycl_b200._lib import check  # noqa: E402

WEAK = {1: (4000, 4000), 2: (4000, 8000), 4: (8000, 8000), 8: (8000, 16000)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=300)
    ap.add_argument("--cells", type=int, nargs=2, default=None)
    ap.add_argument("--fused", type=int, default=1)
    ap.add_argument("--tag", default="run")
    a = ap.parse_args()
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    nx, ny = a.cells if a.cells else WEAK.get(world, (4000, 4000 * world))
    L = lib()
    comms = None
    if world > 1:
        comms = Comms("stamps_%s_%d_%s" % (os.environ.get("MASTER_PORT", "0"), os.getppid(), a.tag), rank, world,
                      device=local)
    s, st = read_config(os.path.join(ROOT, "tests", "decks", "tea_4000_cg.in"), Settings(grid_x_cells=nx, grid_y_cells=ny))
    s.max_iters, s.fuse_p_into_w = a.iters, a.fused
    app = TeaLeaf(s, st, comms, device=local)
    app.solve(0)  # warm-up
    check(L.tl_stamps_enable(app.chunk.handle, a.iters))
    info = app.solve(1)
    n = info.total_iters
    buf = (C.c_ulonglong * (n * 12))()
    check(L.tl_stamps_read(app.chunk.handle, buf, n))
    t = np.ctypeslib.as_array(buf).reshape(n, 3, 4).astype(np.int64)
    fused = bool(L.tl_cg_loop_is_fused(app.chunk.handle, a.fused))
    kernels = [0, 1] if fused else [0, 1, 2]
    names = {0: "calc_pw" if fused else "calc_w", 1: "calc_ur", 2: "calc_p"}
    rows = {}
    sl = slice(5, n - 1)

    def med(x):
        x = x[sl]
        return float(np.median(x)) / 1e3 if len(x) else None

    last = {}
    for kq in kernels:
        t0, t1, t2, t3 = (t[:, kq, j] for j in range(4))
        if kq == 2:  # calc_p has no reduction: only t0 and (multi-rank) t3 are stamped
            last[kq] = np.maximum(t0, t3)
            rows[names[kq]] = {"body_us": med(t3 - t0) if t3.any() else None, "gather_us": None, "halo_us": None}
            continue
        last[kq] = np.maximum(np.maximum(t1, t2), t3)
        rows[names[kq]] = {"body_us": med(t1 - t0), "gather_us": med(t2 - t1) if t2.any() else None,
                           "halo_us": med(t3 - t1) if t3.any() else None}
    for n_, kq in enumerate(kernels):
        nxt = kernels[(n_ + 1) % len(kernels)]
        nt0 = t[:, nxt, 0] if nxt != kernels[0] else np.roll(t[:, nxt, 0], -1)
        rows[names[kq]]["gap_to_next_us"] = med(nt0 - last[kq])
    it_us = med(np.roll(t[:, 0, 0], -1) - t[:, 0, 0])
    out = {"rank": rank, "world": world, "mesh": [nx, ny], "fused": fused, "iters": n, "iteration_us": it_us,
           "solver_ms_per_iter": info.gpu_ms / n, "kernels": rows}
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    path = os.path.join(ROOT, "gpurun_out", "stamps_%s_r%d.json" % (a.tag, rank))
    json.dump(out, open(path, "w"))
    if comms:
        comms.barrier()
    if rank == 0:
        allr = [json.load(open(os.path.join(ROOT, "gpurun_out", "stamps_%s_r%d.json" % (a.tag, r)))) for r in range(world)]
        print("# %s: %d rank(s), mesh %dx%d, %s loop, %d iterations, PDL=%s" % (
            a.tag, world, nx, ny, "fused (calc_pw + calc_ur)" if fused else "three-kernel", n, os.environ.get("TL_PDL", "1")))
        print("# rank  iter_us  ms/iter(events) | " + " | ".join("%s body gather halo gap" % names[q] for q in kernels))
        for o in allr:
            cells = []
            for q in kernels:
                kk = o["kernels"][names[q]]
                cells.append(" ".join("%7.1f" % kk[f] if kk[f] is not None else "      -"
                                      for f in ("body_us", "gather_us", "halo_us", "gap_to_next_us")))
            print("  %d   %8.1f  %8.4f       | %s" % (o["rank"], o["iteration_us"], o["solver_ms_per_iter"], " | ".join(cells)))
        json.dump(allr, open(os.path.join(ROOT, "gpurun_out", "stamps_%s.json" % a.tag), "w"))
    app.close()
    if comms:
        comms.finalise()


if __name__ == "__main__":
    main()
