"""ctypes binding of the CPU oracle (oracle/liboracle.so).  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this module.  The product package (exploringsycl_b200) never does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "liboracle.so")
REF_BIN = os.path.join(HERE, "_ref", "tealeaf_ref")

LEFT, RIGHT, BOTTOM, TOP, EXTERNAL = 0, 1, 2, 3, -1
F_DENSITY, F_ENERGY0, F_ENERGY1, F_U, F_P, F_SD = range(6)
JACOBI, CG, CHEBY, PPCG = range(4)
CONDUCTIVITY, RECIP_CONDUCTIVITY = 1, 2

_dp = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
_ip = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")


def build(force=False):
    """Compile liboracle.so (gcc + OpenMP); also oracle/_ref when /root/reference is present."""
    if force or not os.path.exists(LIB_PATH) or \
            os.path.getmtime(LIB_PATH) < os.path.getmtime(os.path.join(HERE, "tealeaf_oracle.c")):
        subprocess.check_call(["make", "-C", HERE, "liboracle.so"], stdout=subprocess.DEVNULL)
    if os.path.isdir("/root/reference/TeaLeaf"):
        subprocess.check_call(["make", "-C", HERE, "ref"], stdout=subprocess.DEVNULL)


class OrcState(C.Structure):
    _fields_ = [("geometry", C.c_int), ("density", C.c_double), ("energy", C.c_double),
                ("x_min", C.c_double), ("y_min", C.c_double), ("x_max", C.c_double),
                ("y_max", C.c_double), ("radius", C.c_double)]


class OrcDeck(C.Structure):
    _fields_ = [("x_cells", C.c_int), ("y_cells", C.c_int),
                ("xmin", C.c_double), ("ymin", C.c_double), ("xmax", C.c_double), ("ymax", C.c_double),
                ("dt_init", C.c_double), ("end_step", C.c_int), ("max_iters", C.c_int),
                ("eps", C.c_double), ("solver", C.c_int), ("coefficient", C.c_int),
                ("presteps", C.c_int), ("ppcg_inner_steps", C.c_int), ("error_switch", C.c_int),
                ("eps_lim", C.c_double), ("halo_depth", C.c_int), ("summary_frequency", C.c_int),
                ("num_chunks", C.c_int), ("num_states", C.c_int), ("states", OrcState * 16)]


class OrcResult(C.Structure):
    _fields_ = [("iters_a", C.c_int * 64), ("iters_b", C.c_int * 64), ("est_iters", C.c_int * 64),
                ("error", C.c_double * 64), ("eigmin", C.c_double * 64), ("eigmax", C.c_double * 64),
                ("calc_w_calls", C.c_long), ("vol", C.c_double), ("mass", C.c_double),
                ("ie", C.c_double), ("temp", C.c_double), ("wall_solve_s", C.c_double),
                ("cell_iters", C.c_long)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            build()
        L = C.CDLL(LIB_PATH)
        i, d, p = C.c_int, C.c_double, _dp
        pd = C.POINTER(C.c_double)
        sig = {
            "orc_tree_sum": ([p, C.c_long], d),
            "orc_set_sum_mode": ([i], None),
            "orc_set_chunk_data": ([i, i, i, d, d, d, d, p, p, p, p, p], None),
            "orc_set_chunk_initial_state": ([i, i, d, d, p, p], None),
            "orc_set_chunk_state": ([i, i, i, i, d, d, d, d, d, d, d, p, p, p, p, p, p, p], None),
            "orc_store_energy": ([i, i, p, p], None),
            "orc_field_summary": ([i, i, i, p, p, p, p, pd, pd, pd, pd], None),
            "orc_local_halo": ([i, i, i, i, i, p], None),
            "orc_pack": ([i, i, i, i, i, p, p], None),
            "orc_unpack": ([i, i, i, i, i, p, p], None),
            "orc_cg_init": ([i, i, i, i, d, d, p, p, p, p, p, p, p, p, pd], None),
            "orc_cg_calc_w": ([i, i, i, p, p, p, p, pd], None),
            "orc_cg_calc_ur": ([i, i, i, d, p, p, p, p, pd], None),
            "orc_cg_calc_p": ([i, i, i, d, p, p], None),
            "orc_cheby_init": ([i, i, i, d, p, p, p, p, p, p, p], None),
            "orc_cheby_iterate": ([i, i, i, d, d, p, p, p, p, p, p, p], None),
            "orc_ppcg_init": ([i, i, i, d, p, p], None),
            "orc_ppcg_inner_iteration": ([i, i, i, d, d, p, p, p, p, p], None),
            "orc_jacobi_init": ([i, i, i, i, d, d, p, p, p, p, p, p], None),
            "orc_jacobi_iterate": ([i, i, i, p, p, p, p, p, pd], None),
            "orc_copy_u": ([i, i, i, p, p], None),
            "orc_calculate_residual": ([i, i, i, p, p, p, p, p], None),
            "orc_calculate_2norm": ([i, i, i, p, pd], None),
            "orc_finalise": ([i, i, i, p, p, p], None),
            "orc_decompose": ([i, i, i, C.POINTER(i), C.POINTER(i), _ip, _ip, _ip, _ip, _ip], i),
            "orc_eigenvalues": ([p, p, i, pd, pd], i),
            "orc_cheby_coef": ([d, d, i, pd, p, p], None),
            "orc_cheby_est_iterations": ([d, d, d, d], i),
            "orc_deck_defaults": ([C.POINTER(OrcDeck)], None),
            "orc_run_deck": ([C.POINTER(OrcDeck), C.POINTER(OrcResult), C.c_void_p, C.c_void_p], i),
        }
        for name, (args, res) in sig.items():
            f = getattr(L, name)
            f.argtypes = args
            f.restype = res
        _lib = L
    return _lib


# The standard 5-state problem (reference TeaLeaf/tea.in:2-6): (density, energy, xmin, xmax, ymin, ymax)
STANDARD_STATES = [
    (100.0, 0.0001, None),
    (0.1, 25.0, (0.0, 1.0, 1.0, 2.0)),
    (0.1, 0.1, (1.0, 6.0, 1.0, 2.0)),
    (0.1, 0.1, (5.0, 6.0, 1.0, 8.0)),
    (0.1, 0.1, (5.0, 10.0, 7.0, 8.0)),
]


def make_deck(x_cells, y_cells=None, solver=CG, end_step=10, num_chunks=1, max_iters=10000,
              eps=1.0e-15, dt=0.004, states=STANDARD_STATES, **kw):
    d = OrcDeck()
    lib().orc_deck_defaults(C.byref(d))
    d.x_cells = x_cells
    d.y_cells = y_cells or x_cells
    d.xmin, d.ymin, d.xmax, d.ymax = 0.0, 0.0, 10.0, 10.0
    d.dt_init = dt
    d.end_step = end_step
    d.max_iters = max_iters
    d.eps = eps
    d.solver = solver
    d.num_chunks = num_chunks
    d.num_states = len(states)
    for n, (dens, en, box) in enumerate(states):
        s = d.states[n]
        s.geometry, s.density, s.energy = 0, dens, en
        if box:
            s.x_min, s.x_max, s.y_min, s.y_max = box
    for k, v in kw.items():
        setattr(d, k, v)
    return d


def run_deck(deck, want_fields=False, gpu_sum_order=False):
    """gpu_sum_order=True: reductions replay the CUDA backend's tree (bit-for-bit comparable solves)."""
    lib().orc_set_sum_mode(1 if gpu_sum_order else 0)
    try:
        return _run_deck(deck, want_fields)
    finally:
        lib().orc_set_sum_mode(0)


def _run_deck(deck, want_fields=False):
    res = OrcResult()
    u = en = None
    pu = pe = None
    if want_fields:
        u = np.zeros((deck.y_cells, deck.x_cells))
        en = np.zeros((deck.y_cells, deck.x_cells))
        pu, pe = u.ctypes.data, en.ctypes.data
    rc = lib().orc_run_deck(C.byref(deck), C.byref(res), pu, pe)
    if rc:
        raise RuntimeError("orc_run_deck failed rc=%d" % rc)
    n = deck.end_step
    out = dict(iters_a=list(res.iters_a[:n]), iters_b=list(res.iters_b[:n]),
               est_iters=list(res.est_iters[:n]), error=list(res.error[:n]),
               eigmin=list(res.eigmin[:n]), eigmax=list(res.eigmax[:n]),
               calc_w_calls=res.calc_w_calls, vol=res.vol, mass=res.mass, ie=res.ie, temp=res.temp,
               wall_solve_s=res.wall_solve_s, cell_iters=res.cell_iters)
    if want_fields:
        out["u"], out["energy"] = u, en
    return out
