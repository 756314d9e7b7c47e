"""Host-side mirror of the reference TeaLeaf interface over the C-ABI.

Names follow the reference (paths relative to /root/reference/TeaLeaf/):
  Settings / State        settings.h:64-145, defaults settings.h:15-45
  read_config             parse_config.c:14-366 (same keys, same prefix-matching quirks)
  Chunk.run_*             kernel_interface.h:13-71
  Comms                   comms.h:10-20
  decompose_field         initialise.c:34-134
  TeaLeaf.initialise_application / diffuse / solve / field_summary_driver
                          initialise.c:11-31, diffuse.c:10-78, drivers/field_summary_driver.c:8-53
All compute happens in libtealeaf_b200.so on the GPU; this module only sequences calls.
"""
import ctypes as C
import os
from dataclasses import dataclass, field as dc_field
from typing import List, Optional

import numpy as np

from . import _lib
from ._lib import TeaLeafError, TlSolveInfo, TlSolveOpts, TlState, check

# shared.h:33-48
CHUNK_LEFT, CHUNK_RIGHT, CHUNK_BOTTOM, CHUNK_TOP, EXTERNAL_FACE = 0, 1, 2, 3, -1
FIELD_DENSITY, FIELD_ENERGY0, FIELD_ENERGY1, FIELD_U, FIELD_P, FIELD_SD = range(6)
FIELD_U0, FIELD_R, FIELD_W, FIELD_KX, FIELD_KY, FIELD_VOLUME = range(6, 12)
NUM_FIELDS = 6
CONDUCTIVITY, RECIP_CONDUCTIVITY = 1, 2
JACOBI_SOLVER, CG_SOLVER, CHEBY_SOLVER, PPCG_SOLVER = range(4)
RECTANGULAR, CIRCULAR, POINT = range(3)
SOLVER_NAMES = {JACOBI_SOLVER: "Jacobi", CG_SOLVER: "CG", CHEBY_SOLVER: "Chebyshev", PPCG_SOLVER: "PPCG"}


@dataclass
class State:  # settings.h:134-145
    defined: bool = False
    density: float = 0.0
    energy: float = 0.0
    x_min: float = 0.0
    y_min: float = 0.0
    x_max: float = 0.0
    y_max: float = 0.0
    radius: float = 0.0
    geometry: int = RECTANGULAR


@dataclass
class Settings:  # settings.h:64-123 with the defaults of settings.h:15-45
    grid_x_min: float = 0.0
    grid_y_min: float = 0.0
    grid_x_max: float = 100.0
    grid_y_max: float = 100.0
    grid_x_cells: int = 10
    grid_y_cells: int = 10
    dt_init: float = 0.1
    max_iters: int = 10000
    eps: float = 1.0e-15
    end_time: float = 10.0
    end_step: int = 2 ** 31 - 1
    summary_frequency: int = 10
    solver: int = CG_SOLVER
    coefficient: int = CONDUCTIVITY
    error_switch: bool = False
    presteps: int = 30
    eps_lim: float = 1e-5
    check_result: bool = True
    ppcg_inner_steps: int = 10
    preconditioner: bool = False
    num_states: int = 0
    num_chunks: int = 1
    num_chunks_per_rank: int = 1
    num_ranks: int = 1
    halo_depth: int = 2
    rank: int = 0
    dx: float = 0.0
    dy: float = 0.0
    fields_to_exchange: List[bool] = dc_field(default_factory=lambda: [False] * NUM_FIELDS)
    # extensions of this backend (not in the reference)
    profiler_on: bool = False   # deck key the reference ignores (read_config_clean)
    deck_warnings: List[str] = dc_field(default_factory=list)
    batch: int = 0
    fuse_p_into_w: int = 1  # 0: the reference's kernel sequence; non-zero: fused / one-pass kernels (bit-identical), see tl_solve_opts

    def reset_fields_to_exchange(self):  # settings.c:64-70
        self.fields_to_exchange = [False] * NUM_FIELDS

    def solve_opts(self) -> TlSolveOpts:
        o = TlSolveOpts()
        _lib.lib().tl_solve_opts_default(C.byref(o))
        o.solver, o.coefficient, o.max_iters, o.eps = self.solver, self.coefficient, self.max_iters, self.eps
        o.presteps, o.ppcg_inner_steps = self.presteps, self.ppcg_inner_steps
        o.error_switch, o.eps_lim, o.check_result = int(self.error_switch), self.eps_lim, int(self.check_result)
        o.batch = self.batch
        o.fuse_p_into_w = int(self.fuse_p_into_w)
        return o


# ------------------------------------------------------------------------------------------------
# parse_config.c
# ------------------------------------------------------------------------------------------------
def _starts_with(word, line):  # parse_config.c:289-313 (leading whitespace skipped)
    return line.lstrip().startswith(word)


def _read_value(line, word):  # parse_config.c:316-340: first alnum-led token after the key
    pos = line.find(word)
    if pos < 0:
        raise TeaLeafError("Failed to find a value for key '%s'" % word)
    rest = line[pos + len(word):]
    for n, ch in enumerate(rest):
        if ch.isalnum():
            return rest[n:].split()[0]
    raise TeaLeafError("Failed to find a value for key '%s'" % word)


def _atof(tok):  # C atof: longest numeric prefix, 0.0 if none
    import re
    m = re.match(r"[+-]?(\d+\.?\d*([eE][+-]?\d+)?|\.\d+([eE][+-]?\d+)?)", tok)
    return float(m.group(0)) if m else 0.0


def _atoi(tok):
    import re
    m = re.match(r"[+-]?\d+", tok)
    return int(m.group(0)) if m else 0


def read_config(path, settings: Optional[Settings] = None, hygiene: bool = False):
    """parse_config.c:14-74: returns (settings, states). States' extents are shrunk by dx/100
    (parse_config.c:253-260).

    hygiene=False (default) reproduces the reference parser, quirks included (tests pin it against the reference's
    own parse_config.c).  hygiene=True is the cleaned-up reader of SURVEY.md 8f-4, see read_config_clean()."""
    if hygiene:
        return read_config_clean(path, settings)
    s = settings or Settings()
    with open(path) as fh:
        lines = fh.readlines()
    dbl = [("initial_timestep", "dt_init"), ("end_time", "end_time")]
    for line in lines:  # read_settings, parse_config.c:77-192 (order of tests matters: epslim before eps)
        def gd(key, attr):
            if _starts_with(key, line):
                setattr(s, attr, _atof(_read_value(line, key)))
                return True
            return False

        def gi(key, attr):
            if _starts_with(key, line):
                setattr(s, attr, _atoi(_read_value(line, key)))
                return True
            return False
        if gd("initial_timestep", "dt_init") or gd("end_time", "end_time") or gi("end_step", "end_step") \
                or gd("xmin", "grid_x_min") or gd("ymin", "grid_y_min") or gd("xmax", "grid_x_max") \
                or gd("ymax", "grid_y_max"):
            continue
        if s.grid_x_cells == 10 and gi("x_cells", "grid_x_cells"):
            continue
        if s.grid_y_cells == 10 and gi("y_cells", "grid_y_cells"):
            continue
        if gi("summary_frequency", "summary_frequency") or gi("presteps", "presteps") \
                or gi("ppcg_inner_steps", "ppcg_inner_steps") or gd("epslim", "eps_lim") \
                or gi("max_iters", "max_iters") or gd("eps", "eps") \
                or gi("num_chunks_per_rank", "num_chunks_per_rank") or gi("halo_depth", "halo_depth"):
            continue
        for key, attr, val in (("check_result", "check_result", True), ("errswitch", "error_switch", True),
                               ("preconditioner_on", "preconditioner", True),
                               ("use_jacobi", "solver", JACOBI_SOLVER), ("use_cg", "solver", CG_SOLVER),
                               ("use_chebyshev", "solver", CHEBY_SOLVER), ("use_ppcg", "solver", PPCG_SOLVER),
                               ("coefficient_density", "coefficient", CONDUCTIVITY),
                               ("coefficient_inverse_density", "coefficient", RECIP_CONDUCTIVITY)):
            if _starts_with(key, line):
                setattr(s, attr, val)
                break
    s.dx = (s.grid_x_max - s.grid_x_min) / float(s.grid_x_cells)
    s.dy = (s.grid_y_max - s.grid_y_min) / float(s.grid_y_cells)
    # read_states, parse_config.c:195-286
    num_states = 0
    for line in lines:
        if _starts_with("state", line):
            num_states = max(num_states, _atoi(_read_value(line, "state")))
    states = [State() for _ in range(num_states)]
    for line in lines:
        if not _starts_with("state", line):
            continue
        n = _atoi(_read_value(line, "state"))
        st = states[n - 1]
        if st.defined:
            raise TeaLeafError("State number %d defined twice." % n)
        st.density = _atof(_read_value(line, "density"))
        st.energy = _atof(_read_value(line, "energy"))
        if n > 1:
            st.x_min = _atof(_read_value(line, "xmin")) + s.dx / 100.0
            st.y_min = _atof(_read_value(line, "ymin")) + s.dy / 100.0
            st.x_max = _atof(_read_value(line, "xmax")) - s.dx / 100.0
            st.y_max = _atof(_read_value(line, "ymax")) - s.dy / 100.0
            geom = _read_value(line, "geometry")
            if geom == "rectangle":
                st.geometry = RECTANGULAR
            elif geom == "circular":
                st.geometry = CIRCULAR
                st.radius = _atof(_read_value(line, "radius"))
            elif geom == "point":
                st.geometry = POINT
        st.defined = True
    s.num_states = num_states
    return s, states


# keys of read_settings (parse_config.c:88-184): name -> (attribute, kind); upstream decks spell the solver keys with
# a tl_ prefix (Benchmarks/tea_bm_1.out:24-26: tl_max_iters, tl_use_cg, tl_eps), which the reference silently ignores
_CLEAN_KEYS = {
    "initial_timestep": ("dt_init", float), "end_time": ("end_time", float), "end_step": ("end_step", int),
    "xmin": ("grid_x_min", float), "ymin": ("grid_y_min", float), "xmax": ("grid_x_max", float),
    "ymax": ("grid_y_max", float), "x_cells": ("grid_x_cells", int), "y_cells": ("grid_y_cells", int),
    "summary_frequency": ("summary_frequency", int), "presteps": ("presteps", int),
    "ppcg_inner_steps": ("ppcg_inner_steps", int), "epslim": ("eps_lim", float), "max_iters": ("max_iters", int),
    "eps": ("eps", float), "num_chunks_per_rank": ("num_chunks_per_rank", int), "halo_depth": ("halo_depth", int),
    "ch_cg_presteps": ("presteps", int), "ch_cg_epslim": ("eps_lim", float),
}
_CLEAN_SWITCHES = {
    "check_result": ("check_result", True), "errswitch": ("error_switch", True),
    "ch_cg_errswitch": ("error_switch", True), "preconditioner_on": ("preconditioner", True),
    "use_jacobi": ("solver", JACOBI_SOLVER), "use_cg": ("solver", CG_SOLVER),
    "use_chebyshev": ("solver", CHEBY_SOLVER), "use_ppcg": ("solver", PPCG_SOLVER),
    "coefficient_density": ("coefficient", CONDUCTIVITY),
    "coefficient_inverse_density": ("coefficient", RECIP_CONDUCTIVITY),
    "profiler_on": ("profiler_on", True), "use_c_kernels": (None, None), "use_fortran_kernels": (None, None),
}
_CLEAN_IGNORED = ("test_problem", "visit_frequency", "tiles_per_task", "use_vector_loops", "reflective_boundary")


def read_config_clean(path, settings: Optional[Settings] = None):
    """The deck reader without the reference's parsing accidents.  Same keys and the same meaning on a
    well-formed deck -- results are identical there -- but:
      * keys are matched exactly, `key=value` or `key value`, so `eps` can no longer swallow an `epslim` line that comes
        in the wrong order, and signed numbers keep their sign (read_value, parse_config.c:316-340, starts at the first
        alphanumeric character: `xmin=-5.0` is read as 5.0 there);
      * x_cells / y_cells are always honoured (the reference reads them only while the field still holds the default
        10, parse_config.c:102-107: a way of letting -x / -y win that misfires for a deck value of 10);
      * the tl_ spelling of upstream decks (tl_max_iters, tl_eps, tl_use_cg, tl_ppcg_inner_steps, tl_ch_cg_presteps,
        ...) and `profiler_on` are understood instead of being dropped silently (parse_config.c:88-184);
      * circular and point states need no xmax / ymax; unknown keys and malformed values are collected in
        settings.deck_warnings instead of being ignored.
    The dx/100 shrink of the state extents (parse_config.c:253-260) is semantics, not an accident: it is kept."""
    s = settings or Settings()
    preset_x, preset_y = s.grid_x_cells != 10, s.grid_y_cells != 10  # set by the caller (command line) before the deck
    warnings = []
    state_lines = []
    with open(path) as fh:
        lines = fh.readlines()
    in_deck = not any(ln.strip().startswith("*tea") for ln in lines)
    for raw in lines:
        line = raw.split("!")[0].strip()
        if line.startswith("*tea"):
            in_deck = True
            continue
        if line.startswith("*endtea"):
            break
        if not in_deck or not line:
            continue
        tok = line.replace("=", " ").split()
        key = tok[0][3:] if tok[0].startswith("tl_") else tok[0]
        if key == "state":
            state_lines.append(line)
        elif key in _CLEAN_KEYS:
            attr, kind = _CLEAN_KEYS[key]
            try:
                val = kind(float(tok[1])) if kind is int else float(tok[1])
            except (IndexError, ValueError):
                warnings.append("no numeric value for '%s': %r" % (tok[0], raw.rstrip()))
                continue
            if (attr == "grid_x_cells" and preset_x) or (attr == "grid_y_cells" and preset_y):
                continue  # the command line wins
            setattr(s, attr, val)
        elif key in _CLEAN_SWITCHES:
            attr, val = _CLEAN_SWITCHES[key]
            if attr:
                setattr(s, attr, val)
        elif key not in _CLEAN_IGNORED:
            warnings.append("unknown key '%s'" % tok[0])
    if s.grid_x_cells <= 0 or s.grid_y_cells <= 0:
        raise TeaLeafError("x_cells and y_cells must be positive (got %d x %d)" % (s.grid_x_cells, s.grid_y_cells))
    s.dx = (s.grid_x_max - s.grid_x_min) / float(s.grid_x_cells)
    s.dy = (s.grid_y_max - s.grid_y_min) / float(s.grid_y_cells)
    parsed = {}
    for line in state_lines:
        tok = line.replace("=", " ").split()
        try:
            n = int(tok[1])
            kv = dict(zip(tok[2::2], tok[3::2]))
            st = State(defined=True, density=float(kv["density"]), energy=float(kv["energy"]))
            if n > 1:
                geom = kv.get("geometry", "rectangle")
                st.geometry = {"rectangle": RECTANGULAR, "circular": CIRCULAR, "circle": CIRCULAR, "point": POINT}[geom]
                st.x_min = float(kv["xmin"]) + s.dx / 100.0
                st.y_min = float(kv["ymin"]) + s.dy / 100.0
                if st.geometry == RECTANGULAR:
                    st.x_max = float(kv["xmax"]) - s.dx / 100.0
                    st.y_max = float(kv["ymax"]) - s.dy / 100.0
                elif st.geometry == CIRCULAR:
                    st.radius = float(kv["radius"])
                    # what the reference holds after parsing xmax=0 / ymax=0 (never read for this geometry)
                    st.x_max, st.y_max = float(kv.get("xmax", 0.0)) - s.dx / 100.0, float(kv.get("ymax", 0.0)) - s.dy / 100.0
                else:
                    st.x_max, st.y_max = float(kv.get("xmax", 0.0)) - s.dx / 100.0, float(kv.get("ymax", 0.0)) - s.dy / 100.0
        except (KeyError, ValueError, IndexError) as e:
            raise TeaLeafError("malformed state line %r (%s)" % (line, e))
        if n in parsed:
            raise TeaLeafError("State number %d defined twice." % n)
        parsed[n] = st
    if not parsed or sorted(parsed) != list(range(1, len(parsed) + 1)):
        raise TeaLeafError("states must be numbered 1..N without gaps (got %s)" % sorted(parsed))
    states = [parsed[n] for n in sorted(parsed)]
    s.num_states = len(states)
    s.deck_warnings = warnings
    return s, states


def settings_overload(settings: Settings, argv, hygiene: bool = False):
    """main.c:61-99.  hygiene=False reproduces it: -solver / --solver / -s take the next word; -x and -y pass the OPTION
    ITSELF to atoi (main.c:78,83), so `-x 4000` sets grid_x_cells to 0.  hygiene=True reads the number that follows."""
    a = list(argv)
    for n, w in enumerate(a):
        if w in ("-solver", "--solver", "-s"):
            if n + 1 == len(a):
                break
            names = {"cg": CG_SOLVER, "cheby": CHEBY_SOLVER, "ppcg": PPCG_SOLVER, "jacobi": JACOBI_SOLVER}
            if a[n + 1] in names:
                settings.solver = names[a[n + 1]]
            elif hygiene:
                raise TeaLeafError("unknown solver '%s' (cg, cheby, ppcg, jacobi)" % a[n + 1])
        elif w in ("-x", "-y"):
            if n + 1 == len(a):
                break
            if hygiene:
                try:
                    v = int(a[n + 1])
                except ValueError:
                    raise TeaLeafError("%s needs a cell count, got '%s'" % (w, a[n + 1]))
                if v <= 0:
                    raise TeaLeafError("%s needs a positive cell count" % w)
            else:
                v = _atoi(w)  # atoi("-x") == 0
            setattr(settings, "grid_x_cells" if w == "-x" else "grid_y_cells", v)
    return settings


def write_to_visit(nx, ny, x_off, y_off, data, name, step, time, directory="."):
    """shared.c:114-150 (write_to_visit, never called by the reference's own drivers): a brick-of-values header
    <name><step>.bov + the raw little-endian doubles <name><step>.dat.  `data` is the nx x ny interior.  Unlike the
    reference, BRICK_ORIGIN carries the chunk offset and VARIABLE the field name (it writes 0. 0. 0. and 'density')."""
    data = np.ascontiguousarray(data, dtype="<f8")
    assert data.shape == (ny, nx)
    bov = os.path.join(directory, "%s%d.bov" % (name, step))
    dat = "%s%d.dat" % (name, step)
    with open(bov, "w") as f:
        f.write("TIME: %.4f\nDATA_FILE: %s\nDATA_SIZE: %d %d 1\nDATA_FORMAT: DOUBLE\nVARIABLE: %s\n"
                "DATA_ENDIAN: LITTLE\nCENTERING: zone\nBRICK_ORIGIN: %d. %d. 0.\nBRICK_SIZE: %d %d 1\n"
                % (time, dat, nx, ny, name, x_off, y_off, nx, ny))
    data.tofile(os.path.join(directory, dat))
    return bov


def get_checking_value(problems_path, settings):  # field_summary_driver.c:56-91
    try:
        with open(problems_path) as fh:
            for line in fh:
                tok = line.split()
                if len(tok) >= 4 and int(tok[0]) == settings.grid_x_cells and \
                        int(tok[1]) == settings.grid_y_cells and int(tok[2]) == settings.end_step:
                    return float(tok[3])
    except OSError:
        pass
    return None


def decompose_field(grid_x_cells, grid_y_cells, num_chunks, chunk):
    """initialise.c:34-134 for one chunk id. Returns dict(nx, ny, left, bottom, neighbours, x_chunks, y_chunks)."""
    nx, ny, left, bottom, xc, yc = (C.c_int() for _ in range(6))
    nb = (C.c_int * 4)()
    check(_lib.lib().tl_decompose(grid_x_cells, grid_y_cells, num_chunks, chunk, C.byref(nx), C.byref(ny),
                                  C.byref(left), C.byref(bottom), nb, C.byref(xc), C.byref(yc)))
    return dict(nx=nx.value, ny=ny.value, left=left.value, bottom=bottom.value, neighbours=list(nb),
                x_chunks=xc.value, y_chunks=yc.value)


# ------------------------------------------------------------------------------------------------
class Comms:
    """comms.h:10-20 over the shared-memory / NVLink-P2P layer of tl_comms.cu."""

    def __init__(self, session, rank, num_ranks, device=0, host_only=False):
        self.handle = C.c_void_p()
        self.rank, self.num_ranks = rank, num_ranks
        check(_lib.lib().tl_comms_create(C.byref(self.handle), str(session).encode(), rank, num_ranks,
                                         device, int(host_only)))

    def barrier(self):
        check(_lib.lib().tl_comms_barrier(self.handle))

    def sum_over_ranks(self, a):
        v = C.c_double(a)
        check(_lib.lib().tl_comms_sum(self.handle, C.byref(v)))
        return v.value

    def min_over_ranks(self, a):
        v = C.c_double(a)
        check(_lib.lib().tl_comms_min(self.handle, C.byref(v)))
        return v.value

    def send_recv_message(self, send_buffer, recv_buffer, neighbour, send_tag, recv_tag):
        check(_lib.lib().tl_comms_send_recv(self.handle, send_buffer, recv_buffer, send_buffer.size, neighbour,
                                            send_tag, recv_tag))

    def post(self, send_buffer, neighbour, send_tag):  # the MPI_Isend half
        check(_lib.lib().tl_comms_post(self.handle, send_buffer, send_buffer.size, neighbour, send_tag))

    def recv(self, recv_buffer, neighbour, recv_tag):  # the MPI_Irecv + wait half
        check(_lib.lib().tl_comms_recv(self.handle, recv_buffer, recv_buffer.size, neighbour, recv_tag))

    def finalise(self):
        if self.handle:
            _lib.lib().tl_comms_destroy(self.handle)
            self.handle = C.c_void_p()


def _flags(fields_to_exchange):
    return (C.c_int * 6)(*[int(bool(f)) for f in fields_to_exchange])


class Chunk:
    """One mesh chunk resident on one GPU (chunk.h:10-79). Methods mirror kernel_interface.h."""

    def __init__(self, nx, ny, halo_depth=2, max_iters=10000, neighbours=(-1, -1, -1, -1), left=0, bottom=0,
                 device=0):
        self.handle = C.c_void_p()
        self.L = _lib.lib()
        check(self.L.tl_chunk_create(C.byref(self.handle), device, nx, ny, halo_depth, max_iters,
                                     (C.c_int * 4)(*neighbours), left, bottom))
        self.nx, self.ny, self.halo_depth = nx, ny, halo_depth
        self.x, self.y = nx + 2 * halo_depth, ny + 2 * halo_depth  # chunk.c:7-8
        self.neighbours = list(neighbours)
        self.left, self.bottom = left, bottom
        self.max_iters = max_iters
        self.theta = 0.0

    def close(self):
        if self.handle:
            self.L.tl_chunk_destroy(self.handle)
            self.handle = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- field I/O (dense row-major (y, x) images, the reference layout) --
    def write(self, field, array):
        a = np.ascontiguousarray(array, dtype=np.float64)
        if a.shape != (self.y, self.x):
            raise TeaLeafError("field image must have shape (y, x) = (%d, %d)" % (self.y, self.x))
        check(self.L.tl_field_write(self.handle, field, a))

    def read(self, field):
        a = np.empty((self.y, self.x), dtype=np.float64)
        check(self.L.tl_field_read(self.handle, field, a))
        return a

    def read_array(self, which):
        n = [self.x, self.y, self.x + 1, self.y + 1][which]
        a = np.empty(n, dtype=np.float64)
        check(self.L.tl_array_read(self.handle, which, a))
        return a

    def coefficient_array(self, name, n):
        p = getattr(self.L, "tl_" + name)(self.handle)
        return np.ctypeslib.as_array(p, shape=(n,))

    # -- kernel_interface.h --
    def run_set_chunk_data(self, settings):
        check(self.L.tl_run_set_chunk_data(self.handle, settings.grid_x_min, settings.grid_y_min,
                                           settings.dx, settings.dy))

    def run_set_chunk_state(self, settings, states):
        arr = (TlState * len(states))()
        for n, s in enumerate(states):
            arr[n] = TlState(s.geometry, s.density, s.energy, s.x_min, s.y_min, s.x_max, s.y_max, s.radius)
        check(self.L.tl_run_set_chunk_state(self.handle, len(states), arr))

    def run_local_halos(self, settings, depth):
        check(self.L.tl_run_local_halos(self.handle, _flags(settings.fields_to_exchange), depth))

    def run_pack_or_unpack(self, depth, face, pack, field, buffer):
        check(self.L.tl_run_pack_or_unpack(self.handle, depth, face, int(pack), field, buffer))

    def run_store_energy(self):
        check(self.L.tl_run_store_energy(self.handle))

    def run_field_summary(self):
        v = [C.c_double() for _ in range(4)]
        check(self.L.tl_run_field_summary(self.handle, *[C.byref(q) for q in v]))
        return tuple(q.value for q in v)  # vol, mass, ie, temp

    def run_cg_init(self, coefficient, rx, ry, rro=0.0):
        v = C.c_double(rro)
        check(self.L.tl_run_cg_init(self.handle, coefficient, rx, ry, C.byref(v)))
        return v.value

    def run_cg_calc_w(self, pw=0.0):
        v = C.c_double(pw)
        check(self.L.tl_run_cg_calc_w(self.handle, C.byref(v)))
        return v.value

    def run_cg_calc_ur(self, alpha):
        v = C.c_double(0.0)
        check(self.L.tl_run_cg_calc_ur(self.handle, alpha, C.byref(v)))
        return v.value

    def run_cg_calc_p(self, beta):
        check(self.L.tl_run_cg_calc_p(self.handle, beta))

    def run_cheby_init(self):
        check(self.L.tl_run_cheby_init(self.handle, self.theta))

    def run_cheby_iterate(self, alpha, beta):
        check(self.L.tl_run_cheby_iterate(self.handle, alpha, beta))

    def run_jacobi_init(self, coefficient, rx, ry):
        check(self.L.tl_run_jacobi_init(self.handle, coefficient, rx, ry))

    def run_jacobi_iterate(self):
        v = C.c_double(0.0)
        check(self.L.tl_run_jacobi_iterate(self.handle, C.byref(v)))
        return v.value

    def run_ppcg_init(self):
        check(self.L.tl_run_ppcg_init(self.handle, self.theta))

    def run_ppcg_inner_iteration(self, alpha, beta):
        check(self.L.tl_run_ppcg_inner_iteration(self.handle, alpha, beta))

    def run_copy_u(self):
        check(self.L.tl_run_copy_u(self.handle))

    def run_calculate_residual(self):
        check(self.L.tl_run_calculate_residual(self.handle))

    def run_calculate_2norm(self, field):
        v = C.c_double(0.0)
        check(self.L.tl_run_calculate_2norm(self.handle, field, C.byref(v)))
        return v.value

    def run_finalise(self):
        check(self.L.tl_run_finalise(self.handle))

    # -- drivers in the library --
    def halo_update(self, comms, fields_to_exchange, depth):
        check(self.L.tl_halo_update(self.handle, comms.handle if comms else None, _flags(fields_to_exchange), depth))

    def solve(self, comms, settings, rx, ry):
        info = TlSolveInfo()
        o = settings.solve_opts()
        check(self.L.tl_solve(self.handle, comms.handle if comms else None, C.byref(o), rx, ry, C.byref(info)))
        return info

    def timestep(self, comms, settings):
        info = TlSolveInfo()
        o = settings.solve_opts()
        check(self.L.tl_timestep(self.handle, comms.handle if comms else None, C.byref(o), settings.dt_init,
                                 settings.dx, settings.dy, C.byref(info)))
        return info

    def field_summary(self, comms):
        v = [C.c_double() for _ in range(4)]
        check(self.L.tl_field_summary(self.handle, comms.handle if comms else None, *[C.byref(q) for q in v]))
        return tuple(q.value for q in v)

    def sync(self):
        check(self.L.tl_chunk_sync(self.handle))


class TeaLeaf:
    """The application flow of main.c:9-59 for this rank's chunk (one chunk per rank)."""

    def __init__(self, settings: Settings, states: List[State], comms: Optional[Comms] = None, device=0,
                 log=None):
        self.settings, self.states, self.comms = settings, states, comms
        self.log = log or (lambda *a: None)
        s = settings
        s.rank = comms.rank if comms else 0
        s.num_ranks = comms.num_ranks if comms else 1
        s.num_chunks = s.num_ranks * s.num_chunks_per_rank  # initialise.c:37-38
        d = decompose_field(s.grid_x_cells, s.grid_y_cells, s.num_chunks, s.rank)
        self.decomposition = d
        self.chunk = Chunk(d["nx"], d["ny"], s.halo_depth, s.max_iters, d["neighbours"], d["left"], d["bottom"],
                           device)
        if comms:
            check(_lib.lib().tl_comms_attach_chunk(comms.handle, self.chunk.handle))
        self.history = []
        self.initialise_application()

    def initialise_application(self):  # initialise.c:11-31
        s, c = self.settings, self.chunk
        c.run_set_chunk_data(s)
        c.run_set_chunk_state(s, self.states)
        s.reset_fields_to_exchange()
        for f in (FIELD_DENSITY, FIELD_ENERGY0, FIELD_ENERGY1):
            s.fields_to_exchange[f] = True
        c.halo_update(self.comms, s.fields_to_exchange, 2)
        c.run_store_energy()

    def solve(self, tt):  # diffuse.c:23-78
        info = self.chunk.timestep(self.comms, self.settings)
        name = SOLVER_NAMES[self.settings.solver]
        if self.settings.solver in (CG_SOLVER, JACOBI_SOLVER):
            self.log("%s: \t\t\t%d iterations" % (name, info.iters_a))
        else:
            self.log("CG: \t\t\t%d iterations" % info.iters_a)
            self.log("%s: \t\t\t%d iterations" % (name, info.iters_b))
        self.history.append(dict(step=tt + 1, iters_a=info.iters_a, iters_b=info.iters_b,
                                 est_iters=info.est_iters, total_iters=info.total_iters, error=info.error,
                                 eigmin=info.eigmin, eigmax=info.eigmax, gpu_ms=info.gpu_ms,
                                 kernel_launches=info.kernel_launches))
        return info

    def diffuse(self):  # diffuse.c:10-20
        for tt in range(self.settings.end_step):
            self.solve(tt)
        return self.field_summary_driver()

    def field_summary_driver(self):  # field_summary_driver.c:8-30 (all four sums are returned)
        vol, mass, ie, temp = self.chunk.field_summary(self.comms)
        return dict(vol=vol, mass=mass, ie=ie, temp=temp)

    def close(self):
        self.chunk.close()
