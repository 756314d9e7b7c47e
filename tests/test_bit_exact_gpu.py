"""Bit-for-bit parity of whole solves.

The CUDA reductions are deterministic: thread-sequential -> warp butterfly -> tile -> 64-tile group ->
total.  The oracle can replay exactly that tree (oracle.run_deck(gpu_sum_order=True)); with the same
summation order and -fmad=false / -ffp-contract=off arithmetic, the GPU solve and the CPU oracle must
agree in EVERY bit: per-step iteration counts, per-step residual norms, the final field and the summary.
(In its default mode the oracle sums in the reference's 64-ary flat-index order, where a count may move
by one: tests/test_solver_gpu.py.)"""
import ctypes as C
import os

import numpy as np
import pytest

from oracle import oracle as O
from tl_testutil import DECKS, HD, dbl, rng_fields, upload

pytestmark = pytest.mark.gpu


@pytest.fixture()
def gpu_order():
    O.lib().orc_set_sum_mode(1)
    yield
    O.lib().orc_set_sum_mode(0)


@pytest.mark.parametrize("nx,ny", [(37, 23), (513, 300), (1000, 700)])
def test_reductions_bit_exact(gpu_order, nx, ny):
    from exploringsycl_b200 import Chunk
    x, y = nx + 2 * HD, ny + 2 * HD
    ch = Chunk(nx, ny, HD, 100)
    f = rng_fields(nx, ny, seed=99)
    upload(ch, f)
    L = O.lib()
    pw = ch.run_cg_calc_w(0.0)
    w = f["w"].copy(); o = dbl()
    L.orc_cg_calc_w(x, y, HD, f["p"], f["kx"], f["ky"], w, C.byref(o))
    assert pw == o.value
    rrn = ch.run_cg_calc_ur(0.37)
    u, r = f["u"].copy(), f["r"].copy(); o = dbl()
    L.orc_cg_calc_ur(x, y, HD, 0.37, f["p"], w, u, r, C.byref(o))
    assert rrn == o.value
    n = ch.run_calculate_2norm(7)
    o = dbl(); L.orc_calculate_2norm(x, y, HD, r, C.byref(o))
    assert n == o.value
    got = ch.run_field_summary()
    v = [dbl() for _ in range(4)]
    L.orc_field_summary(x, y, HD, f["volume"], f["density"], f["energy0"], u, *[C.byref(q) for q in v])
    assert list(got) == [q.value for q in v]
    ch.close()


@pytest.mark.parametrize("mesh,steps", [((250, 250), 10), ((301, 157), 3), ((500, 500), 2)])
@pytest.mark.parametrize("fused", [1, 0])
def test_cg_solve_bit_exact(mesh, steps, fused):
    from exploringsycl_b200 import Settings, TeaLeaf, read_config
    s, states = read_config(os.path.join(DECKS, "tea_250_cg.in"), Settings(grid_x_cells=mesh[0], grid_y_cells=mesh[1]))
    s.end_step = steps
    s.fuse_p_into_w = fused
    app = TeaLeaf(s, states)
    summary = app.diffuse()
    u = app.chunk.read(3)[2:-2, 2:-2]
    en = app.chunk.read(2)[2:-2, 2:-2]
    hist = app.history
    app.close()
    ores = O.run_deck(O.make_deck(mesh[0], mesh[1], end_step=steps), want_fields=True, gpu_sum_order=True)
    assert [h["iters_a"] for h in hist] == ores["iters_a"]
    assert [h["error"] for h in hist] == ores["error"]
    assert np.array_equal(u, ores["u"]) and np.array_equal(en, ores["energy"])
    assert [summary[k] for k in ("vol", "mass", "ie", "temp")] == [ores[k] for k in ("vol", "mass", "ie", "temp")]


@pytest.mark.parametrize("name,sid", [("cheby", O.CHEBY), ("ppcg", O.PPCG)])
def test_cheby_ppcg_solve_bit_exact(name, sid):
    from exploringsycl_b200 import TeaLeaf, read_config
    s, states = read_config(os.path.join(DECKS, "tea_250_%s.in" % name))
    s.end_step = 3
    app = TeaLeaf(s, states)
    summary = app.diffuse()
    u = app.chunk.read(3)[2:-2, 2:-2]
    hist = app.history
    app.close()
    ores = O.run_deck(O.make_deck(250, solver=sid, end_step=3), want_fields=True, gpu_sum_order=True)
    assert [(h["iters_a"], h["iters_b"], h["est_iters"]) for h in hist] == \
        list(zip(ores["iters_a"], ores["iters_b"], ores["est_iters"]))
    assert [h["eigmin"] for h in hist] == ores["eigmin"] and [h["eigmax"] for h in hist] == ores["eigmax"]
    assert np.array_equal(u, ores["u"])
    assert summary["temp"] == ores["temp"]


def test_jacobi_solve_bit_exact():
    from exploringsycl_b200 import TeaLeaf, read_config
    s, states = read_config(os.path.join(DECKS, "tea_10_jacobi.in"))
    app = TeaLeaf(s, states)
    summary = app.diffuse()
    hist = app.history
    app.close()
    ores = O.run_deck(O.make_deck(10, solver=O.JACOBI), gpu_sum_order=True)
    assert [h["iters_a"] for h in hist] == ores["iters_a"]
    assert [h["error"] for h in hist] == ores["error"]
    assert summary["temp"] == ores["temp"]
