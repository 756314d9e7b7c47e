"""Per-kernel parity: every run_* entry point of the C-ABI against the CPU oracle on the same
seeded inputs.  Element-wise results must be BIT-EXACT (the library is built with -fmad=false,
the oracle with -ffp-contract=off); reductions differ only in summation order and must agree
to 1e-13 relative (fp64, <= 1e6 terms)."""
import ctypes as C

import numpy as np
import pytest

from oracle import oracle as O
from tl_testutil import HD, dbl, download, rel, rng_fields, upload

pytestmark = pytest.mark.gpu

SIZES = [(5, 3), (37, 23), (256, 64), (513, 300), (1000, 700)]
RED_TOL = 1e-13


@pytest.fixture(params=SIZES, ids=lambda s: "%dx%d" % s)
def setup(request):
    from exploringsycl_b200 import Chunk
    nx, ny = request.param
    ch = Chunk(nx, ny, HD, 100)
    f = rng_fields(nx, ny, seed=nx * 1000 + ny)
    upload(ch, f)
    yield ch, f, nx + 2 * HD, ny + 2 * HD
    ch.close()


def test_cg_calc_w(setup):
    ch, f, x, y = setup
    pw = ch.run_cg_calc_w(0.5)
    w = f["w"].copy()
    opw = C.c_double(0.5)
    O.lib().orc_cg_calc_w(x, y, HD, f["p"], f["kx"], f["ky"], w, C.byref(opw))
    assert np.array_equal(ch.read(8), w)
    assert rel(pw, opw.value) < RED_TOL


def test_cg_calc_ur(setup):
    ch, f, x, y = setup
    rrn = ch.run_cg_calc_ur(0.37)
    u, r = f["u"].copy(), f["r"].copy()
    orr = dbl()
    O.lib().orc_cg_calc_ur(x, y, HD, 0.37, f["p"], f["w"], u, r, C.byref(orr))
    got = download(ch, ["u", "r"])
    assert np.array_equal(got["u"], u) and np.array_equal(got["r"], r)
    assert rel(rrn, orr.value) < RED_TOL


def test_cg_calc_p(setup):
    ch, f, x, y = setup
    ch.run_cg_calc_p(0.81)
    p = f["p"].copy()
    O.lib().orc_cg_calc_p(x, y, HD, 0.81, f["r"], p)
    assert np.array_equal(ch.read(4), p)


@pytest.mark.parametrize("coef", [1, 2])
def test_cg_init(setup, coef):
    ch, f, x, y = setup
    rro = ch.run_cg_init(coef, 0.7, 1.3, 0.25)
    o = {k: f[k].copy() for k in ("u", "p", "r", "w", "kx", "ky")}
    orro = C.c_double(0.25)
    O.lib().orc_cg_init(x, y, HD, coef, 0.7, 1.3, f["density"], f["energy"], o["u"], o["p"], o["r"], o["w"],
                        o["kx"], o["ky"], C.byref(orro))
    got = download(ch, list(o))
    for k in o:
        assert np.array_equal(got[k], o[k]), k
    assert rel(rro, orro.value) < RED_TOL


def test_cheby(setup):
    ch, f, x, y = setup
    ch.theta = 3.7
    ch.run_cheby_init()
    o = {k: f[k].copy() for k in ("u", "p", "r", "w")}
    O.lib().orc_cheby_init(x, y, HD, 3.7, o["u"], f["u0"], f["kx"], f["ky"], o["p"], o["r"], o["w"])
    got = download(ch, list(o))
    for k in o:
        assert np.array_equal(got[k], o[k]), k
    ch.run_cheby_iterate(0.6, 0.2)
    O.lib().orc_cheby_iterate(x, y, HD, 0.6, 0.2, o["u"], f["u0"], f["kx"], f["ky"], o["p"], o["r"], o["w"])
    got = download(ch, list(o))
    for k in o:
        assert np.array_equal(got[k], o[k]), k


def test_ppcg(setup):
    ch, f, x, y = setup
    ch.theta = 2.9
    ch.run_ppcg_init()
    o = {k: f[k].copy() for k in ("u", "r", "sd")}
    O.lib().orc_ppcg_init(x, y, HD, 2.9, o["r"], o["sd"])
    assert np.array_equal(ch.read(5), o["sd"])
    ch.run_ppcg_inner_iteration(0.4, 0.9)
    O.lib().orc_ppcg_inner_iteration(x, y, HD, 0.4, 0.9, o["u"], o["r"], f["kx"], f["ky"], o["sd"])
    got = download(ch, list(o))
    for k in o:
        assert np.array_equal(got[k], o[k]), k


@pytest.mark.parametrize("coef", [1, 2])
def test_jacobi(setup, coef):
    ch, f, x, y = setup
    ch.run_jacobi_init(coef, 0.7, 1.3)
    o = {k: f[k].copy() for k in ("u", "u0", "kx", "ky", "r")}
    O.lib().orc_jacobi_init(x, y, HD, coef, 0.7, 1.3, f["density"], f["energy"], o["u0"], o["u"], o["kx"], o["ky"])
    got = download(ch, ["u", "u0", "kx", "ky"])
    for k in got:
        assert np.array_equal(got[k], o[k]), k
    err = ch.run_jacobi_iterate()
    oerr = dbl()
    O.lib().orc_jacobi_iterate(x, y, HD, o["u"], o["u0"], o["r"], o["kx"], o["ky"], C.byref(oerr))
    got = download(ch, ["u", "r"])
    assert np.array_equal(got["u"], o["u"]) and np.array_equal(got["r"], o["r"])
    assert rel(err, oerr.value) < RED_TOL


def test_shared_solver_kernels(setup):
    ch, f, x, y = setup
    ch.run_copy_u()
    u0 = f["u0"].copy()
    O.lib().orc_copy_u(x, y, HD, f["u"], u0)
    assert np.array_equal(ch.read(6), u0)
    ch.run_calculate_residual()
    r = f["r"].copy()
    O.lib().orc_calculate_residual(x, y, HD, f["u"], u0, f["kx"], f["ky"], r)
    assert np.array_equal(ch.read(7), r)
    for fid, arr in ((7, r), (6, u0)):
        n = ch.run_calculate_2norm(fid)
        on = dbl()
        O.lib().orc_calculate_2norm(x, y, HD, arr, C.byref(on))
        assert rel(n, on.value) < RED_TOL
    ch.run_finalise()
    en = f["energy"].copy()
    O.lib().orc_finalise(x, y, HD, f["u"], f["density"], en)
    assert np.array_equal(ch.read(2), en)
    ch.run_store_energy()
    assert np.array_equal(ch.read(2), f["energy0"])


def test_field_summary(setup):
    ch, f, x, y = setup
    got = ch.run_field_summary()
    v = [dbl() for _ in range(4)]
    O.lib().orc_field_summary(x, y, HD, f["volume"], f["density"], f["energy0"], f["u"], *[C.byref(q) for q in v])
    for a, b in zip(got, v):
        assert rel(a, b.value) < RED_TOL


def test_reduction_is_deterministic(setup):
    ch, f, x, y = setup
    a = ch.run_cg_calc_w(0.0)
    b = ch.run_cg_calc_w(0.0)
    assert a == b


def test_two_chunks_per_rank_accumulate_semantics():
    """SURVEY.md 8b scalar-return convention with num_chunks_per_rank > 1: the drivers loop over the rank's chunks with
    ONE accumulator; run_cg_init / run_cg_calc_w ADD their chunk's sum to it (cg.cpp:133,194), run_cg_calc_ur,
    run_calculate_2norm, run_jacobi_iterate and run_field_summary ASSIGN (cg.cpp:253, solver_methods.cpp:116,
    jacobi.cpp:116, field_summary.cpp:147-150), so the last chunk's value survives -- as in the reference."""
    from exploringsycl_b200 import Chunk
    chunks, fields = [], []
    for n, (nx, ny) in enumerate([(70, 40), (33, 57)]):
        ch = Chunk(nx, ny, HD, 100)
        f = rng_fields(nx, ny, seed=100 + n)
        upload(ch, f)
        chunks.append(ch)
        fields.append(f)
    L = O.lib()
    own_pw, own_rro, own_rrn = [], [], []
    for ch, f in zip(chunks, fields):
        x, y = ch.x, ch.y
        o = dbl(); w = f["w"].copy()
        L.orc_cg_calc_w(x, y, HD, f["p"], f["kx"], f["ky"], w, C.byref(o))
        own_pw.append(o.value)
    # cg_driver.c:72-85: pw = 0; for each chunk run_cg_calc_w(chunk, settings, &pw)
    pw = 0.0
    for ch in chunks:
        pw = ch.run_cg_calc_w(pw)
    assert rel(pw, own_pw[0] + own_pw[1]) < RED_TOL
    # run_cg_calc_ur ASSIGNS: called with the same variable for both chunks, the second chunk's value is what is left
    rrn = [ch.run_cg_calc_ur(0.37) for ch in chunks]
    for ch, f, got in zip(chunks, fields, rrn):
        x, y = ch.x, ch.y
        u, r = f["u"].copy(), f["r"].copy()
        w = f["w"].copy(); o = dbl()
        L.orc_cg_calc_w(x, y, HD, f["p"], f["kx"], f["ky"], w, C.byref(o))
        o2 = dbl()
        L.orc_cg_calc_ur(x, y, HD, 0.37, f["p"], w, u, r, C.byref(o2))
        assert rel(got, o2.value) < RED_TOL
    # run_cg_init accumulates into *rro
    rro = 0.125
    for ch in chunks:
        rro = ch.run_cg_init(1, 0.7, 1.3, rro)
    exp = C.c_double(0.125)
    for ch, f in zip(chunks, fields):
        o = {k: f[k].copy() for k in ("u", "p", "r", "w", "kx", "ky")}
        L.orc_cg_init(ch.x, ch.y, HD, 1, 0.7, 1.3, f["density"], f["energy"], o["u"], o["p"], o["r"], o["w"], o["kx"],
                      o["ky"], C.byref(exp))
    assert rel(rro, exp.value) < RED_TOL
    for ch in chunks:
        ch.close()
