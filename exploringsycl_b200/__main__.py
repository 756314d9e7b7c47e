"""python -m exploringsycl_b200 -- the application flow of the reference's main.c:9-59 on the B200 backend, with the
command line and deck hygiene of SURVEY.md 8f-4.

    python -m exploringsycl_b200 [--deck tea.in] [-x N] [-y N] [-s cg|cheby|ppcg|jacobi] [--visit] [--json tea.json]
                                 [--problems tea.problems] [--device D] [--reference-quirks]
    python -m torch.distributed.run --nproc-per-node N -m exploringsycl_b200 ...     (one rank per GPU)

What differs from the reference host (all of it opt-out with --reference-quirks, which reproduces main.c / parse_config.c
to the letter -- that behaviour is what tests/test_host_logic.py pins against the reference's own parser):
  * -x / -y work (main.c:76-85 passes the option string itself to atoi, so they set the mesh to 0 cells);
  * the deck is read by read_config_clean(): exact keys, signed numbers, tl_ aliases of upstream decks, profiler_on,
    warnings for unknown keys, x_cells / y_cells always honoured;
  * all four field-summary sums, per-step solver rates and a JSON sidecar are reported, and --visit
    writes <field><step>.bov/.dat bricks for density, energy and temperature after the last step (shared.c:114-150).
The printed reference lines ("CG: n iterations", "Expected / Actual", PASSED / FAILED) keep the reference's format.
"""
import argparse
import json
import os
import sys
import time

from . import (Comms, Settings, TeaLeaf, TeaLeafError, get_checking_value, read_config, settings_overload,
               write_to_visit)
from .tealeaf import CG_SOLVER, CHEBY_SOLVER, JACOBI_SOLVER, PPCG_SOLVER, SOLVER_NAMES

BYTES = {JACOBI_SOLVER: 56.0, CG_SOLVER: 104.0, CHEBY_SOLVER: 88.0}  # algorithmic B/cell/iteration, SURVEY.md 8d


def step_work(s, info):
    """cell-iterations and algorithmic bytes of one solve (same accounting as c_kernels/cuda/diffuse_overload.cpp)"""
    cells = float(s.grid_x_cells) * float(s.grid_y_cells)
    iters = float(info.total_iters)
    if s.solver == PPCG_SOLVER:
        iters += float(info.iters_b) * s.ppcg_inner_steps
        bpc = 104.0 * (info.total_iters - info.iters_b) + (128.0 + 80.0 * s.ppcg_inner_steps) * info.iters_b
    elif s.solver == CHEBY_SOLVER:
        bpc = 104.0 * (info.total_iters - info.iters_b) + 88.0 * info.iters_b
    else:
        bpc = BYTES[s.solver] * info.total_iters
    return cells * iters, cells * bpc


def main(argv=None):
    ap = argparse.ArgumentParser(prog="python -m exploringsycl_b200", add_help=True)
    ap.add_argument("--deck", default="tea.in")
    ap.add_argument("--problems", default="tea.problems")
    ap.add_argument("-x", type=str, default=None, dest="x")
    ap.add_argument("-y", type=str, default=None, dest="y")
    ap.add_argument("-s", "-solver", "--solver", default=None, dest="solver")
    ap.add_argument("--visit", action="store_true")
    ap.add_argument("--json", default="tea.json")
    ap.add_argument("--device", type=int, default=None)
    ap.add_argument("--reference-quirks", action="store_true")
    a = ap.parse_args(argv)
    hygiene = not a.reference_quirks
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    device = a.device if a.device is not None else int(os.environ.get("LOCAL_RANK", "0"))
    out = (lambda *w: print(*w, flush=True)) if rank == 0 else (lambda *w: None)

    s = Settings()
    over = []
    for opt, val in (("-x", a.x), ("-y", a.y), ("-s", a.solver)):
        if val is not None:
            over += [opt, val]
    try:
        if hygiene:
            # the command line is applied BEFORE the deck is read and the mesh is built; the deck's x_cells / y_cells /
            # use_* lines yield to it
            settings_overload(s, ["tealeaf"] + over, hygiene=True)
            solver_cli = s.solver if a.solver is not None else None
            s, states = read_config(a.deck, s, hygiene=True)
            if solver_cli is not None:
                s.solver = solver_cli
        else:
            s, states = read_config(a.deck, s)
    except (TeaLeafError, OSError, ZeroDivisionError) as e:
        out("tealeaf: %s" % e)
        return 2
    for w in s.deck_warnings:
        out("WARNING (deck): %s" % w)
    comms = None
    if world > 1:
        comms = Comms("cli_%s_%d" % (os.environ.get("MASTER_PORT", "0"), os.getppid()), rank, world, device=device)
    out("Solution Parameters:\n\tx_cells = %d\n\ty_cells = %d\n\tsolver = %s\n\tranks = %d" % (
        s.grid_x_cells, s.grid_y_cells, SOLVER_NAMES[s.solver], world))
    app = TeaLeaf(s, states, comms, device=device)
    if not hygiene:
        # main.c:25-29: the reference applies the command line AFTER initialise_application has read the deck and built
        # the chunks: only the solver choice still matters; -x / -y overwrite the cell counts (with atoi of the option
        # string: 0) that the checking-value lookup reads later
        settings_overload(s, ["tealeaf"] + over, hygiene=False)
    steps, t0 = [], time.perf_counter()
    for tt in range(s.end_step):
        out("\nTimestep %d" % (tt + 1))
        info = app.solve(tt)
        if s.solver in (CG_SOLVER, JACOBI_SOLVER):
            out("%s: \t\t\t%d iterations" % (SOLVER_NAMES[s.solver], info.iters_a))
        else:
            out("CG: \t\t\t%d iterations" % info.iters_a)
            out("%s: \t\t\t%d iterations" % ("Cheby" if s.solver == CHEBY_SOLVER else "PPCG", info.iters_b))
        ci, by = step_work(s, info)
        rate = ci / (info.gpu_ms * 1e-3) if info.gpu_ms > 0 else 0.0
        gbs = by / 1e9 / (info.gpu_ms * 1e-3) if info.gpu_ms > 0 else 0.0
        steps.append(dict(step=tt + 1, iters_a=info.iters_a, iters_b=info.iters_b, est_iters=info.est_iters,
                          total_iters=info.total_iters, error=info.error, solver_ms=info.gpu_ms,
                          cell_iters_per_s=rate, algorithmic_gb_per_s=gbs, kernel_launches=info.kernel_launches))
        if tt % s.summary_frequency == 0:
            sm = app.field_summary_driver()
            out("Field summary: \t\tvol %.15e mass %.15e ie %.15e temp %.15e" % (sm["vol"], sm["mass"], sm["ie"], sm["temp"]))
        out("Wallclock: \t\t%.3fs" % (time.perf_counter() - t0))
        out("Error: \t\t\t%.6e" % info.error)
        out("Solver rate: \t\t%d iterations in %.3f ms, %.4e cell-iterations/s, %.1f GB/s (algorithmic)" % (
            info.total_iters, info.gpu_ms, rate, gbs))
        if s.profiler_on:
            out("Profiler: \t\t%d kernel launches, %.3f ms on the solver stream" % (info.kernel_launches, info.gpu_ms))
    sm = app.field_summary_driver()
    out("Field summary: \t\tvol %.15e mass %.15e ie %.15e temp %.15e" % (sm["vol"], sm["mass"], sm["ie"], sm["temp"]))
    rc = 0
    if s.check_result:  # field_summary_driver.c:32-52
        out("\nChecking results...")
        expect = get_checking_value(a.problems, s)
        if expect is None:
            out("\nWARNING: Problem was not found in the test problems file.")
            expect = 1.0
        out("Expected %.15e" % expect)
        out("Actual   %.15e" % sm["temp"])
        qa = abs(100.0 * (sm["temp"] / expect) - 100.0)
        out("This run %s (Difference is within %.8f%%)" % ("PASSED" if qa < 0.001 else "FAILED", qa))
    if a.visit:
        c, hd = app.chunk, s.halo_depth
        sl = (slice(hd, -hd), slice(hd, -hd))
        d = app.decomposition
        for name, fid in (("density", 0), ("energy", 2), ("temperature", 3)):
            tag = name if world == 1 else "%s.r%d." % (name, rank)
            write_to_visit(c.nx, c.ny, d["left"], d["bottom"], c.read(fid)[sl], tag, s.end_step, s.end_step * s.dt_init)
    if rank == 0 and a.json:
        tot_ms = sum(q["solver_ms"] for q in steps)
        side = dict(backend="exploringsycl_b200", solver=SOLVER_NAMES[s.solver].lower(), grid=[s.grid_x_cells, s.grid_y_cells],
                    ranks=world, end_step=s.end_step, eps=s.eps, max_iters=s.max_iters, steps=steps,
                    total_iters=sum(q["total_iters"] for q in steps), solver_ms=tot_ms,
                    wallclock_s=time.perf_counter() - t0, deck_warnings=list(s.deck_warnings),
                    field_summary=dict(volume=sm["vol"], mass=sm["mass"], internal_energy=sm["ie"], temperature=sm["temp"]))
        with open(a.json, "w") as f:
            json.dump(side, f, indent=1)
    app.close()
    if comms:
        comms.finalise()
    return rc


if __name__ == "__main__":
    sys.exit(main())
