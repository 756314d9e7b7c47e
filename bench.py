#!/usr/bin/env python
"""bench.py -- TeaLeaf CG hot path on B200: cell-iterations/s and HBM roofline.

    python bench.py --gpus N --steps K --warmup W [--impl reference]
    (N > 1: launched by the driver through torch.distributed.run, one rank per GPU)

A "step" is one TeaLeaf timestep (diffuse.c:23-78): depth-2 halo of energy+density, the CG solve to
eps = 1e-15 (cg_driver.c), residual check, finalise, energy halo -- on the standard 5-state deck
(reference TeaLeaf/Benchmarks/tea_bm_5.in geometry), synthetic by construction (no input files).

  N = 1 : 4000 x 4000 mesh  (BASELINE.json configs[1], tea_bm_5)
  N > 1 : weak scaling at 16 M cells per GPU:
          4000x8000 @2, 8000x8000 @4, 8000x16000 @8, decomposed by the reference's decompose_field; halos and the
          two per-iteration scalars move GPU to GPU over NVLink peer memory inside the solver kernels (no NCCL on
          the data path; torch.distributed/NCCL only provides the barrier and the max-over-ranks of the timing).

metric  = CG cell-iterations per second = x_cells * y_cells * (CG iterations executed) / time
value   = whole job, state resident in HBM, timed with CUDA events on the solver stream, max over ranks
e2e     = same metric through tl_timestep_host(): density + energy images uploaded from pinned host
          memory and energy + summary read back inside the timed region, every step
roofline= dominant kernel of the iteration (cg_calc_pw or cg_calc_ur, 48 B/cell each) timed live with CUDA events on
          the launching stream vs MEASURED_PEAKS.json; traffic = ncu DRAM bytes per launch (profiles/traffic.json)
cpu_baseline / --impl reference = the CPU oracle (C + OpenMP restatement of the reference kernels and
          drivers; the reference's own SYCL kernels cannot be built here) on the box's host cores.
"""
import argparse
import ctypes as C
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

BYTES_PER_CELL_ITER = 104  # calc_w 32 + calc_ur 48 + calc_p 24
KERNELS = ((0, "cg_calc_w", 32), (1, "cg_calc_ur", 48), (2, "cg_calc_p", 24), (3, "cg_calc_pw", 48))
WEAK_MESH = {1: (4000, 4000), 2: (4000, 8000), 4: (8000, 8000), 8: (8000, 16000)}


def mesh_for(n_gpus):
    if n_gpus in WEAK_MESH:
        return WEAK_MESH[n_gpus]
    return (4000, 4000 * n_gpus)


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, burst copy)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx, self.rows, self.proc = gpu_index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([t.strip() for t in line.split(",")])

    def stop(self):
        if self.proc:
            self.proc.terminate()
        sm = [float(r[1]) for r in self.rows if len(r) >= 9 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 9 for i in range(4) if r[5 + i] == "Active"})
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


def deck_settings(nx, ny, max_iters=10000):
    from exploringsycl_b200 import Settings, read_config
    s, states = read_config(os.path.join(ROOT, "tests", "decks", "tea_4000_cg.in"),
                            Settings(grid_x_cells=nx, grid_y_cells=ny))
    s.max_iters = max_iters
    s.end_step = 1 << 30
    return s, states


def cpu_port_run(nx, ny, iters_cap, steps):
    """Times the CPU oracle (C + OpenMP) on `steps` timesteps capped at iters_cap CG iterations."""
    from oracle import oracle as O
    d = O.make_deck(nx, ny, end_step=steps, max_iters=iters_cap)
    t0 = time.perf_counter()
    r = O.run_deck(d)
    wall = time.perf_counter() - t0
    return r, wall


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path on the host cores. The
    reference's own kernels are SYCL (unbuildable here); the arm runs the C/OpenMP port of
    them (oracle/) under the reference's driver call order, all host threads, bounded sample."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    nx, ny = args.mesh if args.mesh else mesh_for(args.gpus)
    cores = os.cpu_count() or 1
    os.environ.setdefault("OMP_NUM_THREADS", str(cores))
    cap = args.ref_iters
    # one untimed probe step sizes the sample so the whole run stays within a few minutes
    r, wall = cpu_port_run(nx, ny, 4, 1)
    per_iter = max(wall / 6.0, 1e-4)
    budget = 150.0 / max(args.steps + args.warmup, 1)
    cap = max(4, min(cap, int(budget / per_iter)))
    for _ in range(args.warmup):
        cpu_port_run(nx, ny, cap, 1)
    t0 = time.perf_counter()
    cell_iters = 0
    for _ in range(args.steps):
        r, _w = cpu_port_run(nx, ny, cap, 1)
        cell_iters += r["cell_iters"]
    wall = time.perf_counter() - t0
    value = cell_iters / wall
    sample = "%dx%d mesh, %d timestep(s) of the standard deck from t=0, CG capped at %d iterations each" % (
        nx, ny, args.steps, cap)
    line = {"impl": "reference", "metric": "CG cell-iterations/s", "value": value, "unit": "cell-iter/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * wall / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "TeaLeaf CG fp64 %dx%d (tea_bm_5 deck geometry), CPU bounded sample" % (nx, ny),
                       "sample": sample},
            "cpu_baseline": {"value": value, "unit": "cell-iter/s", "cores": cores, "kind": "port",
                             "sample": sample},
            "e2e": {"value": value, "unit": "cell-iter/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gb_per_s_algorithmic": value * BYTES_PER_CELL_ITER / 1e9}
    print(json.dumps(line))
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--ref-iters", type=int, default=200)
    ap.add_argument("--cpu-seconds", type=float, default=15.0)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--mesh", type=int, nargs=2, default=None)
    ap.add_argument("--solver", default="cg", choices=["cg", "cheby", "ppcg", "jacobi"],
                    help="non-default solvers are reported under their own metric name (BASELINE configs[3])")
    ap.add_argument("--max-iters", type=int, default=10000)
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    n_gpus = max(args.gpus, world)
    # stdout carries exactly ONE JSON line: anything libraries print (e.g. NCCL's version banner) goes to
    # stderr until the line is written
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    import torch
    import torch.distributed as dist
    from exploringsycl_b200 import Comms, TeaLeaf, lib
    from exploringsycl_b200._lib import TlSolveInfo, check

    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"  # keep stdout to the one JSON line
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(v):
        if world == 1:
            return v
        t = torch.tensor([v], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(v):
        if world == 1:
            return v
        t = torch.tensor([v], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    nx, ny = args.mesh if args.mesh else mesh_for(n_gpus)
    L = lib()
    comms = None
    if world > 1:
        # all workers of one torchrun launch share the agent as parent: a per-launch name, so a segment left
        # behind by a crashed earlier run on the same port can never be picked up
        session = "bench_%s_%d" % (os.environ.get("MASTER_PORT", "0"), os.getppid())
        comms = Comms(session, rank, world, device=local_rank)
    s, states = deck_settings(nx, ny, args.max_iters)
    s.solver = {"jacobi": 0, "cg": 1, "cheby": 2, "ppcg": 3}[args.solver]
    inner = s.ppcg_inner_steps if args.solver == "ppcg" else 0

    def matvecs(info):  # solver-loop iterations (+ the reduction-free inner steps of PPCG)
        return info.total_iters + info.iters_b * inner
    if os.environ.get("TL_BENCH_FUSED") is not None:
        s.fuse_p_into_w = int(os.environ["TL_BENCH_FUSED"])
    app = TeaLeaf(s, states, comms, device=local_rank)
    ch = app.chunk
    cells = nx * ny

    # ---------------- value: state resident in HBM ----------------
    for tt in range(args.warmup):
        app.solve(tt)
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    l0 = L.tl_kernel_launch_count()
    check(L.tl_timer_start(ch.handle))
    t0 = time.perf_counter()
    iters = 0
    per_step = []
    for tt in range(args.steps):
        info = app.solve(args.warmup + tt)
        iters += matvecs(info)
        per_step.append(info.iters_a if args.solver in ("cg", "jacobi") else [info.iters_a, info.iters_b])
    ms = C.c_double()
    check(L.tl_timer_stop(ch.handle, C.byref(ms)))
    barrier()
    wall = time.perf_counter() - t0
    clocks = sampler.stop()
    launches = L.tl_kernel_launch_count() - l0
    gpu_s = max_over_ranks(ms.value / 1e3)
    value = cells * iters / gpu_s
    summary = app.field_summary_driver()

    # ---------------- e2e: host buffers, copies inside the timed region ----------------
    e2e = None
    if not args.no_e2e:
        import numpy as np
        nbytes = ch.x * ch.y * 8
        hd_p, he_p = L.tl_host_alloc_pinned(nbytes), L.tl_host_alloc_pinned(nbytes)
        dens = np.ctypeslib.as_array(C.cast(hd_p, C.POINTER(C.c_double)), shape=(ch.y, ch.x))
        ener = np.ctypeslib.as_array(C.cast(he_p, C.POINTER(C.c_double)), shape=(ch.y, ch.x))
        check(L.tl_field_read(ch.handle, 0, dens))
        check(L.tl_field_read(ch.handle, 2, ener))
        o = s.solve_opts()
        info = TlSolveInfo()
        summ = (C.c_double * 4)()
        barrier()
        t0 = time.perf_counter()
        e_iters = 0
        for tt in range(args.steps):
            check(L.tl_timestep_host(ch.handle, comms.handle if comms else None, C.byref(o), s.dt_init, s.dx,
                                     s.dy, hd_p, he_p, C.byref(info), C.byref(summ)))
            e_iters += matvecs(info)
        barrier()
        e_wall = max_over_ranks(time.perf_counter() - t0)
        e2e = {"value": cells * e_iters / e_wall, "unit": "cell-iter/s",
               "h2d_bytes_per_step": int(sum_over_ranks(2 * nbytes)),
               "d2h_bytes_per_step": int(sum_over_ranks(nbytes + 32)),
               "ms_per_step": 1e3 * e_wall / args.steps, "iters": e_iters,
               "temp": summ[3]}
        L.tl_host_free_pinned(hd_p)
        L.tl_host_free_pinned(he_p)

    # ---------------- roofline: the hot kernels timed live (rank 0's chunk) ----------------
    peak, peak_src = measured_peak()
    kern = {}
    chunk_cells = ch.nx * ch.ny
    for which, name, bpc in KERNELS:
        kms = C.c_double()
        check(L.tl_time_kernel(ch.handle, which, 40, C.byref(kms)))
        kern[name] = {"ms": kms.value, "bytes_per_cell": bpc, "gb_s": chunk_cells * bpc / kms.value / 1e6,
                      "frac": chunk_cells * bpc / kms.value / 1e6 / peak}
    traffic = None
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tp):
        try:
            traffic = json.load(open(tp)).get("cg_calc_ur", {}).get("dram_bytes_per_launch")
        except Exception:
            traffic = None
    # the resident CG iteration is cg_calc_pw (fused p-update + matvec) + cg_calc_ur: the dominant kernel
    # is whichever of the two takes longer per launch
    lr_nb = app.decomposition["x_chunks"] > 1
    fused_run = s.fuse_p_into_w == 2 or (s.fuse_p_into_w == 1 and not lr_nb)
    dom_name = "cg_calc_pw" if (fused_run and kern["cg_calc_pw"]["ms"] >= kern["cg_calc_ur"]["ms"]) \
        else "cg_calc_ur"
    dom = kern[dom_name]
    if os.path.exists(tp):
        try:
            traffic = json.load(open(tp)).get(dom_name, {}).get("dram_bytes_per_launch")
        except Exception:
            traffic = None
    roofline = {"bound": "hbm", "kernel": dom_name, "achieved": dom["gb_s"], "peak": peak, "unit": "GB/s",
                "frac": dom["frac"], "traffic": traffic, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": chunk_cells * dom["bytes_per_cell"], "kernels": kern,
                "solve_gb_s_104B": value / n_gpus * BYTES_PER_CELL_ITER / 1e9,
                "solve_frac_104B": value / n_gpus * BYTES_PER_CELL_ITER / 1e9 / peak}

    # ---------------- CPU baseline (rank 0, N = 1): the oracle port on the host cores ----------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        cores = os.cpu_count() or 1
        os.environ.setdefault("OMP_NUM_THREADS", str(cores))
        r, w = cpu_port_run(nx, ny, 4, 1)
        per_iter = max(w / 6.0, 1e-4)
        cap = max(4, min(2000, int(args.cpu_seconds / per_iter)))
        r, w = cpu_port_run(nx, ny, cap, 1)
        cpu = {"value": r["cell_iters"] / w, "unit": "cell-iter/s", "cores": cores, "kind": "port",
               "sample": "%dx%d mesh, first timestep of the same deck, CG capped at %d iterations, %.1f s" % (
                   nx, ny, cap, w)}

    if rank == 0:
        metric = {"cg": "CG", "cheby": "Chebyshev", "ppcg": "PPCG (outer+inner)", "jacobi": "Jacobi"}[args.solver]
        line = {"metric": metric + " cell-iterations/s", "value": value, "unit": "cell-iter/s", "n_gpus": n_gpus,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * gpu_s / args.steps,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
                "data": "synthetic",
                "config": {"workload": "TeaLeaf %s fp64 %dx%d, standard 5-state deck (tea_bm_5 geometry), "
                                       "eps 1e-15, one timestep per step, %s" % (
                                           metric, nx, ny, "1 GPU" if n_gpus == 1 else
                                           "%d GPUs %dx%d chunks" % (n_gpus, app.decomposition["x_chunks"],
                                                                     app.decomposition["y_chunks"])),
                           "cells_per_gpu": cells // n_gpus, "cg_iterations_per_step": per_step,
                           "l2": "inputs larger than L2: 7 live fields x %.0f MB per GPU" % (ch.x * ch.y * 8 / 1e6),
                           "bytes_per_cell_iter": BYTES_PER_CELL_ITER,
                           "iteration": "cg_calc_pw + cg_calc_ur (96 B/cell moved)"
                           if fused_run
                           else "cg_calc_w + cg_calc_ur + cg_calc_p (104 B/cell moved)"},
                "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline,
                "cpu_baseline": cpu, "wall_s": wall, "summary": summary}
        sys.stdout.flush()
        os.dup2(real_stdout, 1)
        print(json.dumps(line), flush=True)
        os.dup2(2, 1)
    app.close()
    if comms:
        comms.finalise()
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
