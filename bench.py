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
parity  = BEFORE anything is timed: a 250x250 deck decomposed over the N ranks (three-kernel and fused resident
          loops) must equal the CPU oracle's N-chunk run in every bit (iteration counts, per-step residuals, the four
          field-summary sums), and 60 back-to-back halo exchanges verified cell by cell on the device; the run
          FAILS otherwise.  (The oracle is the checker here, never the thing measured.)
roofline= dominant kernel of the iteration (cg_calc_pw or cg_calc_ur, 48 B/cell each) timed live with CUDA events on
          the launching stream vs MEASURED_PEAKS.json; traffic = DRAM bytes per launch of that kernel from the
          committed ncu capture (profiles/traffic.json; null when the capture does not cover the kernel); in_loop = the
          same kernel's duration inside the resident loop, from the device stamps of iteration_profile
iteration_profile = where an iteration's time goes: %globaltimer stamps taken by the loop kernels themselves (body,
          NVLink all-gather of the partial sums, halo hand-shake, launch gap), median over 300 iterations, min / max over
          the ranks
loop_form_tuning = (N >= 4) the two bit-identical forms of the CG iteration (fused two-kernel / three-kernel) timed on a
          capped solve before the timed region; the faster one runs
extra   = (N > 1) the other named configs of BASELINE.json, device-timed the same way with the iteration count capped:
          strong scaling of 4000x4000 over the N GPUs; at N = 8 also CG 16000^2 and 32000^2 and Chebyshev / PPCG 8000^2
cpu_baseline / --impl reference = the CPU oracle (C + OpenMP restatement of the reference kernels and
          drivers; the reference's own SYCL kernels cannot be built here) on the box's host cores: solver time only
          (the oracle's wall_solve_s), CG capped at >= 100 iterations per step whatever the mesh.
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
CPU_MIN_ITERS = 100   # floor of the CPU sample per step, whatever the mesh
CPU_BUDGET_S = 180.0  # target wall time of a whole --impl reference run


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


# ------------------------------------------------------------------------------------------------
# CPU arm: the oracle port (C + OpenMP), solver time only
# ------------------------------------------------------------------------------------------------
def cpu_port_run(nx, ny, iters_cap, steps):
    """`steps` timesteps of the standard deck from t = 0 with CG capped at iters_cap iterations each.
    Returns (cell_iters, solver seconds): the oracle times cg_driver only (wall_solve_s), so deck
    allocation and set-up -- which the GPU arm does not time either -- stay outside."""
    from oracle import oracle as O
    d = O.make_deck(nx, ny, end_step=steps, max_iters=iters_cap)
    r = O.run_deck(d)
    return r["cell_iters"], r["wall_solve_s"]


def cpu_iters_cap(nx, ny, seconds_per_step, hard_max=2000):
    """Iterations per step that fit `seconds_per_step` on this host, never below CPU_MIN_ITERS."""
    ci, w = cpu_port_run(nx, ny, 8, 1)  # probe (also the first-touch / thread-pool warm-up)
    per_iter = max(w / max(ci / (nx * ny), 1.0), 1e-6)
    return max(CPU_MIN_ITERS, min(hard_max, int(seconds_per_step / per_iter))), per_iter


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
    threads = int(os.environ["OMP_NUM_THREADS"])
    cap, per_iter = cpu_iters_cap(nx, ny, CPU_BUDGET_S / max(args.steps + args.warmup, 1))
    if args.ref_iters:
        cap = max(CPU_MIN_ITERS if nx * ny > 100 * 100 else 1, min(cap, args.ref_iters))
    if args.warmup:
        cpu_port_run(nx, ny, cap, args.warmup)
    cell_iters, solve_s = cpu_port_run(nx, ny, cap, args.steps)
    value = cell_iters / solve_s
    sample = ("%dx%d mesh, %d timestep(s) of the standard deck from t=0, CG capped at %d iterations each "
              "(floor %d), solver time only, %d OpenMP threads" % (nx, ny, args.steps, cap, CPU_MIN_ITERS, threads))
    line = {"impl": "reference", "metric": "CG cell-iterations/s", "value": value, "unit": "cell-iter/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * solve_s / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "TeaLeaf CG fp64 %dx%d (tea_bm_5 deck geometry), CPU bounded sample" % (nx, ny),
                       "sample": sample, "iters_cap": cap},
            "cpu_baseline": {"value": value, "unit": "cell-iter/s", "cores": threads, "kind": "port",
                             "sample": sample, "iters_cap": cap, "host_cpus": cores},
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
    ap.add_argument("--ref-iters", type=int, default=0, help="upper bound of the CPU sample (0 = sized by time)")
    ap.add_argument("--cpu-seconds", type=float, default=15.0)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="skip the other named configs at N > 1")
    ap.add_argument("--no-parity", action="store_true")
    ap.add_argument("--no-profile", action="store_true", help="skip the stamped solve behind iteration_profile")
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
    from exploringsycl_b200 import Comms, TeaLeaf, lib, read_config
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

    def min_over_ranks(v):
        return -max_over_ranks(-v)

    def sum_over_ranks(v):
        if world == 1:
            return v
        t = torch.tensor([v], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    L = lib()
    n_sessions = [0]

    def new_comms():
        # all workers of one torchrun launch share the agent as parent: a per-launch name, so a segment left
        # behind by a crashed earlier run on the same port can never be picked up.  One endpoint per chunk.
        if world == 1:
            return None
        n_sessions[0] += 1
        session = "bench_%s_%d_%d" % (os.environ.get("MASTER_PORT", "0"), os.getppid(), n_sessions[0])
        return Comms(session, rank, world, device=local_rank)

    solver_id = {"jacobi": 0, "cg": 1, "cheby": 2, "ppcg": 3}

    # ---------------- parity: before anything is timed ----------------
    parity = None
    if not args.no_parity:
        parity = parity_leg(L, new_comms, world, rank, local_rank, sum_over_ranks)
        if not (parity["n_chunk_bit_exact"] and parity["halo_bit_exact"] in (True, None)):
            raise SystemExit("bench.py: parity against the CPU oracle FAILED: %s" % json.dumps(parity))

    def timed_case(nx, ny, solver, max_iters, steps, warmup, fuse=None, keep=False):
        """One configuration, state resident in HBM: `steps` timesteps timed with CUDA events, max over ranks."""
        comms = new_comms()
        s, states = deck_settings(nx, ny, max_iters)
        s.solver = solver_id[solver]
        if fuse is not None:
            s.fuse_p_into_w = fuse
        inner = s.ppcg_inner_steps if solver == "ppcg" else 0
        app = TeaLeaf(s, states, comms, device=local_rank)
        for tt in range(warmup):
            app.solve(tt)
        barrier()
        l0 = L.tl_kernel_launch_count()
        check(L.tl_timer_start(app.chunk.handle))
        iters, per_step = 0, []
        for tt in range(steps):
            info = app.solve(warmup + tt)
            iters += info.total_iters + info.iters_b * inner  # solver-loop iterations (+ PPCG's inner steps)
            per_step.append(info.iters_a if solver in ("cg", "jacobi") else [info.iters_a, info.iters_b])
        ms = C.c_double()
        check(L.tl_timer_stop(app.chunk.handle, C.byref(ms)))
        barrier()
        gpu_s = max_over_ranks(ms.value / 1e3)
        out = {"value": nx * ny * iters / gpu_s, "gpu_s": gpu_s, "iters": iters, "per_step": per_step,
               "launches": L.tl_kernel_launch_count() - l0, "summary": app.field_summary_driver(),
               "decomposition": [app.decomposition["x_chunks"], app.decomposition["y_chunks"]]}
        if keep:
            out.update(app=app, comms=comms, settings=s)
        else:
            app.close()
            if comms:
                comms.finalise()
        return out

    nx, ny = args.mesh if args.mesh else mesh_for(n_gpus)
    fuse_env = int(os.environ["TL_BENCH_FUSED"]) if os.environ.get("TL_BENCH_FUSED") is not None else None
    cells = nx * ny

    # ---------------- loop form: two kernels (fused p-update + matvec) or three, whichever is faster here ----------------
    # Both forms give bit-identical results (parity leg above).  With left/right neighbours the better one depends on the
    # chunk shape, so at N >= 4 both are timed on a capped solve BEFORE the timed region and the faster one is used;
    # every rank takes the decision from the same max-over-ranks times.
    loop_tune = None
    if world >= 4 and fuse_env is None and args.solver == "cg":
        t3 = timed_case(nx, ny, "cg", 300, 1, 1, 0)
        t2 = timed_case(nx, ny, "cg", 300, 1, 1, 2)
        fuse_env = 2 if t2["gpu_s"] <= t3["gpu_s"] else 0
        loop_tune = {"three_kernel_ms_per_iter": 1e3 * t3["gpu_s"] / max(t3["iters"], 1),
                     "fused_ms_per_iter": 1e3 * t2["gpu_s"] / max(t2["iters"], 1),
                     "chosen": "fused" if fuse_env == 2 else "three_kernel"}

    # ---------------- value: state resident in HBM ----------------
    sampler = ClockSampler(local_rank)
    sampler.start()
    t0 = time.perf_counter()
    main_case = timed_case(nx, ny, args.solver, args.max_iters, args.steps, args.warmup, fuse_env, keep=True)
    wall = time.perf_counter() - t0
    clocks = sampler.stop()
    app, comms, s = main_case["app"], main_case["comms"], main_case["settings"]
    ch = app.chunk
    value, gpu_s, launches = main_case["value"], main_case["gpu_s"], main_case["launches"]
    inner = s.ppcg_inner_steps if args.solver == "ppcg" else 0

    # ---------------- where an iteration's time goes: %globaltimer stamps of the loop kernels (all ranks) ----------------
    iteration_profile = None
    if args.solver == "cg" and not args.no_profile:
        iteration_profile = stamp_profile(L, app, s, world, max_over_ranks, min_over_ranks)

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
            e_iters += info.total_iters + info.iters_b * inner
        barrier()
        e_wall = max_over_ranks(time.perf_counter() - t0)
        e2e = {"value": cells * e_iters / e_wall, "unit": "cell-iter/s",
               "h2d_bytes_per_step": int(sum_over_ranks(2 * nbytes)),
               "d2h_bytes_per_step": int(sum_over_ranks(nbytes + 32)),
               "ms_per_step": 1e3 * e_wall / args.steps, "iters": e_iters,
               "temp": summ[3]}
        L.tl_host_free_pinned(hd_p)
        L.tl_host_free_pinned(he_p)

    # ---------------- roofline: the hot kernels timed live (this rank's chunk; rank 0 reports) ----------------
    peak, peak_src = measured_peak()
    kern = {}
    chunk_cells = ch.nx * ch.ny
    for which, name, bpc in KERNELS:
        kms = C.c_double()
        check(L.tl_time_kernel(ch.handle, which, 40, C.byref(kms)))
        kern[name] = {"ms": kms.value, "bytes_per_cell": bpc, "gb_s": chunk_cells * bpc / kms.value / 1e6,
                      "frac": chunk_cells * bpc / kms.value / 1e6 / peak}
    # the resident CG iteration is cg_calc_pw (fused p-update + matvec) + cg_calc_ur, or the three reference kernels:
    # the dominant kernel is whichever of the loop's kernels takes longest per launch
    fused_run = bool(L.tl_cg_loop_is_fused(ch.handle, s.fuse_p_into_w))
    dom_name = "cg_calc_pw" if (fused_run and kern["cg_calc_pw"]["ms"] >= kern["cg_calc_ur"]["ms"]) \
        else "cg_calc_ur"
    dom = kern[dom_name]
    traffic, traffic_src = None, None
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tp) and (nx, ny, n_gpus) == (4000, 4000, 1):
        try:
            ent = json.load(open(tp)).get(dom_name, {})
            traffic, traffic_src = ent.get("dram_bytes_per_launch"), ent.get("source")
        except Exception:
            traffic = None
    in_loop = None
    if iteration_profile and (dom_name + ".body") in iteration_profile:
        b = iteration_profile[dom_name + ".body"]
        us = b["max"] if isinstance(b, dict) else b
        if us > 0:
            in_loop = {"us": us, "gb_s": chunk_cells * dom["bytes_per_cell"] / us / 1e3,
                       "frac": chunk_cells * dom["bytes_per_cell"] / us / 1e3 / peak,
                       "how": "device %globaltimer stamps inside the resident loop (slowest rank)"}
    roofline = {"bound": "hbm", "kernel": dom_name, "achieved": dom["gb_s"], "peak": peak, "unit": "GB/s",
                "frac": dom["frac"], "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": chunk_cells * dom["bytes_per_cell"], "kernels": kern, "in_loop": in_loop,
                "solve_gb_s_104B": value / n_gpus * BYTES_PER_CELL_ITER / 1e9,
                "solve_frac_104B": value / n_gpus * BYTES_PER_CELL_ITER / 1e9 / peak}
    decomposition = app.decomposition
    app.close()
    if comms:
        comms.finalise()

    # ---------------- the other named configs (BASELINE.json configs[2..4]), N > 1 ----------------
    extra = None
    if world > 1 and not args.no_extra and args.solver == "cg" and not args.mesh:
        extra = {}
        cases = [("cg_4000_strong", 4000, 4000, "cg", 1000)]
        if world == 8:
            cases += [("cg_16000", 16000, 16000, "cg", 400), ("cg_32000", 32000, 32000, "cg", 200),
                      # The CG pre-steps only hand over once the residual has dropped below 1 (cheby_driver.c:30-32): ~1900
                      # iterations into the solve at this size (profiles/cheby_ppcg_r02.txt).  The caps leave ~2100
                      # Chebyshev iterations / ~600 PPCG outer steps (6000 inner steps) in the timed solve; run to
                      # convergence the same solves take 30000+ iterations (6 s Chebyshev, 33 s PPCG per timestep).
                      ("cheby_8000", 8000, 8000, "cheby", 4000), ("ppcg_8000", 8000, 8000, "ppcg", 2500)]
        for name, ex, ey, solver, cap in cases:
            try:
                r = timed_case(ex, ey, solver, cap, 2 if solver == "cg" else 1, 1, fuse_env)
                it_ms = 1e3 * r["gpu_s"] / max(r["iters"], 1)
                bpc = {"cg": 104, "cheby": 88, "ppcg": 80}[solver]  # SURVEY.md 8d algorithmic bytes per cell-iteration
                extra[name] = {"mesh": [ex, ey], "solver": solver, "value": r["value"], "unit": "cell-iter/s",
                               "max_iters": cap, "steps": 2 if solver == "cg" else 1, "warmup": 1, "iters": r["iters"],
                               "iterations_per_step": r["per_step"],
                               "ms_per_iter": it_ms, "decomposition": r["decomposition"],
                               "gb_s_per_gpu_algorithmic": r["value"] / n_gpus * bpc / 1e9,
                               "frac_of_peak_algorithmic": r["value"] / n_gpus * bpc / 1e9 / peak,
                               "temp": r["summary"]["temp"]}
            except Exception as e:  # an extra must never cost the headline line
                extra[name] = {"error": str(e)[:300]}

    # ---------------- CPU baseline (rank 0, N = 1): the oracle port on the host cores ----------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        cores = os.cpu_count() or 1
        os.environ.setdefault("OMP_NUM_THREADS", str(cores))
        threads = int(os.environ["OMP_NUM_THREADS"])
        cap, _ = cpu_iters_cap(nx, ny, args.cpu_seconds)
        ci, w = cpu_port_run(nx, ny, cap, 1)
        cpu = {"value": ci / w, "unit": "cell-iter/s", "cores": threads, "kind": "port", "host_cpus": cores,
               "iters_cap": cap,
               "sample": "%dx%d mesh, first timestep of the same deck, CG capped at %d iterations, solver time only "
                         "%.1f s, %d OpenMP threads" % (nx, ny, cap, w, threads)}

    if rank == 0:
        metric = {"cg": "CG", "cheby": "Chebyshev", "ppcg": "PPCG (outer+inner)", "jacobi": "Jacobi"}[args.solver]
        line = {"metric": metric + " cell-iterations/s", "value": value, "unit": "cell-iter/s", "n_gpus": n_gpus,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * gpu_s / args.steps,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
                "data": "synthetic",
                "config": {"workload": "TeaLeaf %s fp64 %dx%d, standard 5-state deck (tea_bm_5 geometry), "
                                       "eps 1e-15, one timestep per step, %s" % (
                                           metric, nx, ny, "1 GPU" if n_gpus == 1 else
                                           "%d GPUs %dx%d chunks" % (n_gpus, decomposition["x_chunks"],
                                                                     decomposition["y_chunks"])),
                           "cells_per_gpu": cells // n_gpus, "cg_iterations_per_step": main_case["per_step"],
                           "l2": "inputs larger than L2: 7 live fields x %.0f MB per GPU" % (ch.x * ch.y * 8 / 1e6),
                           "bytes_per_cell_iter": BYTES_PER_CELL_ITER,
                           "iteration": "cg_calc_pw + cg_calc_ur (96 B/cell moved)"
                           if fused_run
                           else "cg_calc_w + cg_calc_ur + cg_calc_p (104 B/cell moved)"},
                "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline,
                "cpu_baseline": cpu, "parity": parity, "iteration_profile": iteration_profile, "loop_form_tuning": loop_tune,
                "extra": extra, "wall_s": wall,
                "summary": main_case["summary"]}
        sys.stdout.flush()
        os.dup2(real_stdout, 1)
        print(json.dumps(line), flush=True)
        os.dup2(2, 1)
    if world > 1:
        dist.destroy_process_group()
    return 0


def stamp_profile(L, app, s, world, max_over_ranks, min_over_ranks, iters=300):
    """One more solve of the benchmark problem, capped at `iters` iterations, with the loop kernels' %globaltimer stamps
    switched on (tl_stamps_enable, see tools/stamps.py).  Medians over the iterations, in microseconds; for N > 1 the
    minimum and maximum over the ranks (clocks are per GPU: only same-rank differences are taken).
      body   first CTA past its dependency wait -> tail CTA holds this rank's partial sum (streaming + grid reduction)
      gather tail CTA: NVLink all-gather of the ranks' partials (latency on the last rank to arrive, + skew on the others)
      halo   tail CTA: halo hand-shake with the neighbours
      gap    end of the kernel's tail -> first CTA of the next kernel past its wait"""
    import numpy as np
    from exploringsycl_b200._lib import check
    ch = app.chunk
    keep = s.max_iters
    s.max_iters = iters
    check(L.tl_stamps_enable(ch.handle, iters))
    info = app.solve(0)
    n = info.total_iters
    buf = (C.c_ulonglong * (n * 12))()
    check(L.tl_stamps_read(ch.handle, buf, n))
    check(L.tl_stamps_enable(ch.handle, 0))
    s.max_iters = keep
    t = np.ctypeslib.as_array(buf).reshape(n, 3, 4).astype(np.int64)
    fused = bool(L.tl_cg_loop_is_fused(ch.handle, s.fuse_p_into_w))
    kernels = [0, 1] if fused else [0, 1, 2]
    names = {0: "cg_calc_pw" if fused else "cg_calc_w", 1: "cg_calc_ur", 2: "cg_calc_p"}
    sl = slice(5, n - 1)

    def med(x):
        return float(np.median(x[sl])) / 1e3 if n > 8 else 0.0

    out = {"iterations": n, "loop": "fused" if fused else "three_kernel", "unit": "us (median over iterations)"}
    last = {}
    vals = {}
    for kq in kernels:
        t0, t1, t2, t3 = (t[:, kq, j] for j in range(4))
        if kq == 2:
            last[kq] = np.maximum(t0, t3)
            vals[names[kq] + ".body_incl_handshake"] = med(t3 - t0) if t3.any() else 0.0
            continue
        last[kq] = np.maximum(np.maximum(t1, t2), t3)
        vals[names[kq] + ".body"] = med(t1 - t0)
        vals[names[kq] + ".gather"] = med(t2 - t1)
        if t3.any():
            vals[names[kq] + ".halo"] = med(t3 - t1)
    for i_, kq in enumerate(kernels):
        nxt = kernels[(i_ + 1) % len(kernels)]
        nt0 = t[:, nxt, 0] if nxt != kernels[0] else np.roll(t[:, nxt, 0], -1)
        vals[names[kq] + ".gap_to_next"] = med(nt0 - last[kq])
    vals["iteration"] = med(np.roll(t[:, 0, 0], -1) - t[:, 0, 0])
    for k_, v in vals.items():  # same keys in the same order on every rank
        out[k_] = v if world == 1 else {"min": min_over_ranks(v), "max": max_over_ranks(v)}
    return out


def parity_leg(L, new_comms, world, rank, local_rank, sum_over_ranks):
    """Decomposed 250x250 CG deck (3 timesteps; three-kernel and fused resident loops) vs the CPU oracle's
    N-chunk run with the GPU summation order: every rank must hold the oracle's iteration counts, per-step
    residuals and the four summary sums bit for bit.  Then 60 back-to-back halo exchanges of changing field
    sets / depths, generated and verified cell by cell on the device (tl_halo_stress)."""
    from exploringsycl_b200 import TeaLeaf, read_config
    from exploringsycl_b200._lib import check
    from oracle import oracle as O
    os.environ.setdefault("OMP_NUM_THREADS", str(max(1, (os.cpu_count() or 1) // max(world, 1))))
    ores = O.run_deck(O.make_deck(250, end_step=3, num_chunks=world), gpu_sum_order=True)
    bad, halo_bad = 0, None
    detail = {}
    for label, fuse in (("three_kernel", 0), ("fused", 2)):
        s, states = read_config(os.path.join(ROOT, "tests", "decks", "tea_250_cg.in"))
        s.end_step, s.fuse_p_into_w = 3, fuse
        comms = new_comms()
        app = TeaLeaf(s, states, comms, device=local_rank)
        summary = app.diffuse()
        ok = ([h["iters_a"] for h in app.history] == ores["iters_a"]
              and [h["error"] for h in app.history] == ores["error"]
              and all(summary[k] == ores[k] for k in ("vol", "mass", "ie", "temp")))
        bad += 0 if ok else 1
        detail[label] = {"iters": [h["iters_a"] for h in app.history], "temp": summary["temp"]}
        if world > 1 and fuse == 2:
            n = C.c_long(-1)
            check(L.tl_halo_stress(app.chunk.handle, comms.handle, 250, 250, 60, 2, C.byref(n)))
            halo_bad = n.value
        app.close()
        if comms:
            comms.finalise()
    bad_all = int(sum_over_ranks(float(bad)))
    halo_all = None if halo_bad is None else int(sum_over_ranks(float(halo_bad)))
    return {"n_chunk_bit_exact": bad_all == 0, "halo_bit_exact": None if halo_all is None else halo_all == 0,
            "chunks": world, "oracle_iters": ores["iters_a"], "oracle_temp": ores["temp"], "gpu": detail,
            "what": "tea_250_cg deck, 3 timesteps, decomposed over %d rank(s): CG iteration counts, per-step rrn and "
                    "vol/mass/ie/temp equal the CPU oracle's %d-chunk run bit for bit (three-kernel and fused loops); "
                    "60 back-to-back halo exchanges verified on the device" % (world, world)}


if __name__ == "__main__":
    sys.exit(main())
