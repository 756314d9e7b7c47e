"""Whole-solve parity against the oracle's run of the same deck (reference call order).

Bar (BASELINE.json north_star): iteration counts agree and the field summary (vol, mass, ie, temp)
agrees within 1e-10 relative in fp64.  Reduction order differs between the GPU (tile tree) and the
oracle (reference 64-ary tree), so a CG count may move by +-1 when sqrt(rrn) grazes eps
(SURVEY.md section 4); the tests allow +-1 per step and report exact matches."""
import os

import numpy as np
import pytest

from oracle import oracle as O
from tl_testutil import DECKS, GOLDEN, rel

pytestmark = pytest.mark.gpu
SUMMARY_TOL = 1e-10


def run_gpu(deck_file, **over):
    from exploringsycl_b200 import TeaLeaf, read_config
    s, states = read_config(os.path.join(DECKS, deck_file))
    for k, v in over.items():
        setattr(s, k, v)
    app = TeaLeaf(s, states)
    summary = app.diffuse()
    u = app.chunk.read(3)
    hist = app.history
    app.close()
    return s, summary, hist, u


def check_summary(summary, ores):
    for k in ("vol", "mass", "ie", "temp"):
        assert rel(summary[k], ores[k]) < SUMMARY_TOL, (k, summary[k], ores[k])


@pytest.mark.parametrize("n", [10, 250, 500])
def test_cg_deck(n):
    s, summary, hist, u = run_gpu("tea_%d_cg.in" % n)
    ores = O.run_deck(O.make_deck(n), want_fields=True)
    got = [h["iters_a"] for h in hist]
    assert all(abs(a - b) <= 1 for a, b in zip(got, ores["iters_a"])), (got, ores["iters_a"])
    check_summary(summary, ores)
    from exploringsycl_b200 import get_checking_value
    exp = get_checking_value(os.path.join(GOLDEN, "tea_problems.txt"), s)
    assert abs(100.0 * (summary["temp"] / exp) - 100.0) < 0.001  # field_summary_driver.c:42-43
    hd = s.halo_depth
    ui = u[hd:-hd, hd:-hd]
    assert np.max(np.abs(ui - ores["u"])) / np.max(np.abs(ores["u"])) < 1e-10


def test_cg_nonsquare_odd():
    """Non-square, odd-sized mesh: exercises the ragged last column pair and partial tiles."""
    from exploringsycl_b200 import Settings, TeaLeaf, read_config
    # x_cells / y_cells are only parsed while still at the default 10 (parse_config.c:102-107)
    s2, states = read_config(os.path.join(DECKS, "tea_250_cg.in"), Settings(grid_x_cells=301, grid_y_cells=157))
    s2.end_step = 2
    app = TeaLeaf(s2, states)
    summary = app.diffuse()
    hist = app.history
    app.close()
    ores = O.run_deck(O.make_deck(301, 157, end_step=2))
    got = [h["iters_a"] for h in hist]
    assert all(abs(a - b) <= 1 for a, b in zip(got, ores["iters_a"])), (got, ores["iters_a"])
    check_summary(summary, ores)


def test_cg_max_iters_cap():
    s, summary, hist, u = run_gpu("tea_250_cg.in", max_iters=50, end_step=1)
    ores = O.run_deck(O.make_deck(250, end_step=1, max_iters=50))
    assert hist[0]["iters_a"] == 50 == ores["iters_a"][0]
    check_summary(summary, ores)


@pytest.mark.parametrize("solver,name", [(O.CHEBY, "cheby"), (O.PPCG, "ppcg")])
def test_cheby_ppcg_deck(solver, name):
    s, summary, hist, u = run_gpu("tea_250_%s.in" % name)
    ores = O.run_deck(O.make_deck(250, solver=solver))
    assert [h["iters_a"] for h in hist] == ores["iters_a"]
    gb, ob = [h["iters_b"] for h in hist], ores["iters_b"]
    # the 2-norm is only sampled every 10th Chebyshev iteration: counts can move by one sample
    step = 10 if solver == O.CHEBY else 1
    assert all(abs(a - b) <= step for a, b in zip(gb, ob)), (gb, ob)
    assert rel(hist[0]["eigmin"], ores["eigmin"][0]) < 1e-9 and rel(hist[0]["eigmax"], ores["eigmax"][0]) < 1e-9
    for k in ("vol", "mass", "ie"):
        assert rel(summary[k], ores[k]) < SUMMARY_TOL
    # these solvers stop at |rrn| < eps (not sqrt): the answer itself is only converged to ~1e-9
    tol = SUMMARY_TOL if gb == ob else 5e-9
    assert rel(summary["temp"], ores["temp"]) < tol


def test_small_mesh_never_leaves_cg_phase():
    """10x10: Chebyshev/PPCG decks converge inside the CG pre-steps (|rrn| < eps end test)."""
    for solver, sid in (("use_chebyshev", O.CHEBY), ("use_ppcg", O.PPCG)):
        from exploringsycl_b200 import TeaLeaf, read_config
        s, states = read_config(os.path.join(DECKS, "tea_10_cg.in"))
        s.solver = sid
        app = TeaLeaf(s, states)
        summary = app.diffuse()
        hist = app.history
        app.close()
        ores = O.run_deck(O.make_deck(10, solver=sid))
        assert [h["iters_b"] for h in hist] == ores["iters_b"] == [0] * 10
        assert all(abs(a - b) <= 1 for a, b in zip([h["iters_a"] for h in hist], ores["iters_a"]))
        check_summary(summary, ores)


def test_jacobi_deck():
    s, summary, hist, u = run_gpu("tea_10_jacobi.in")
    ores = O.run_deck(O.make_deck(10, solver=O.JACOBI))
    assert [h["iters_a"] for h in hist] == ores["iters_a"]
    check_summary(summary, ores)


def test_plugin_api_cg_iteration_matches_resident_loop():
    """Host-driven run_cg_* sequence (cg_driver.c call order) == the device-resident loop, bit for bit."""
    from exploringsycl_b200 import TeaLeaf, read_config
    s, states = read_config(os.path.join(DECKS, "tea_250_cg.in"))
    s.end_step = 1
    a = TeaLeaf(s, states)
    a.diffuse()
    s2, states2 = read_config(os.path.join(DECKS, "tea_250_cg.in"))
    b = TeaLeaf(s2, states2)
    c = b.chunk
    rx = s2.dt_init / (s2.dx * s2.dx)
    ry = s2.dt_init / (s2.dy * s2.dy)
    s2.reset_fields_to_exchange(); s2.fields_to_exchange[2] = s2.fields_to_exchange[0] = True
    c.halo_update(None, s2.fields_to_exchange, 2)
    rro = c.run_cg_init(s2.coefficient, rx, ry, 0.0)
    s2.reset_fields_to_exchange(); s2.fields_to_exchange[3] = s2.fields_to_exchange[4] = True
    c.halo_update(None, s2.fields_to_exchange, 1)
    c.run_copy_u()
    tt = 0
    for tt in range(s2.max_iters):
        pw = c.run_cg_calc_w(0.0)
        alpha = rro / pw
        rrn = c.run_cg_calc_ur(alpha)
        beta = rrn / rro
        c.run_cg_calc_p(beta)
        rro = rrn
        c.halo_update(None, s2.fields_to_exchange, 1)
        if abs(rrn) ** 0.5 < s2.eps:
            break
    assert tt == a.history[0]["iters_a"]
    assert np.array_equal(a.chunk.read(3)[2:-2, 2:-2], c.read(3)[2:-2, 2:-2])
    a.close(); b.close()


@pytest.mark.parametrize("mesh", [(250, 250), (301, 157), (64, 513)])
def test_fused_p_into_w_is_bit_identical(mesh):
    """The fused p-update + matvec kernel (48 B/cell instead of 24 + 32) must reproduce the three-kernel
    iteration bit for bit: same counts, same u / p / r / w fields."""
    from exploringsycl_b200 import Settings, TeaLeaf, read_config
    res = []
    for fused in (False, True):
        s, states = read_config(os.path.join(DECKS, "tea_250_cg.in"),
                                Settings(grid_x_cells=mesh[0], grid_y_cells=mesh[1]))
        s.end_step = 2
        s.fuse_p_into_w = fused
        app = TeaLeaf(s, states)
        summary = app.diffuse()
        res.append((summary, [h["iters_a"] for h in app.history],
                    {f: app.chunk.read(f) for f in (3, 4, 7, 8, 2)}))
        app.close()
    assert res[0][1] == res[1][1]
    assert res[0][0] == res[1][0]
    hd = 2
    for f in res[0][2]:
        a, b = res[0][2][f], res[1][2][f]
        assert np.array_equal(a[hd - 1:-(hd - 1), hd - 1:-(hd - 1)], b[hd - 1:-(hd - 1), hd - 1:-(hd - 1)]), f


@pytest.mark.parametrize("stages", [3, 4, 8])
def test_tma_row_pipeline_variant_is_bit_identical(stages):
    """tl_set_pw_pipeline(1, ..): the fused p+w kernel on a cp.async.bulk / mbarrier shared-memory row pipeline with
    persistent CTAs (tl_bulk.cu) must give the register kernel's results bit for bit (fields, counts, residuals)."""
    from exploringsycl_b200 import Settings, TeaLeaf, lib, read_config
    L = lib()
    res = []
    try:
        for mode in (0, 1):
            assert L.tl_set_pw_pipeline(mode, stages, 0) == 0
            s, states = read_config(os.path.join(DECKS, "tea_250_cg.in"), Settings(grid_x_cells=301, grid_y_cells=157))
            s.end_step = 2
            app = TeaLeaf(s, states)
            summary = app.diffuse()
            res.append((summary, [(h["iters_a"], h["error"]) for h in app.history],
                        {f: app.chunk.read(f)[2:-2, 2:-2] for f in (3, 7, 8, 2)}))
            app.close()
    finally:
        L.tl_set_pw_pipeline(0, 4, 0)
    assert res[0][0] == res[1][0] and res[0][1] == res[1][1]
    for f in res[0][2]:
        assert np.array_equal(res[0][2][f], res[1][2][f]), f


@pytest.mark.parametrize("rows,batch", [(8, 1), (16, 4), (32, 2)])
def test_tuning_does_not_change_fields(rows, batch):
    """Load-batch depth never changes results; rows per tile only reorders the reduction."""
    from exploringsycl_b200 import TeaLeaf, lib, read_config
    s, states = read_config(os.path.join(DECKS, "tea_250_cg.in"))
    s.end_step = 1
    base = TeaLeaf(s, states)
    base.diffuse()
    try:
        for k in range(4):
            assert lib().tl_set_tuning(k, rows, batch) == 0
        s2, states2 = read_config(os.path.join(DECKS, "tea_250_cg.in"))
        s2.end_step = 1
        app = TeaLeaf(s2, states2)
        app.diffuse()
        assert abs(app.history[0]["iters_a"] - base.history[0]["iters_a"]) <= 1
        u0, u1 = base.chunk.read(3)[2:-2, 2:-2], app.chunk.read(3)[2:-2, 2:-2]
        assert np.max(np.abs(u0 - u1)) / np.max(np.abs(u0)) < 1e-12
        app.close()
    finally:
        for k, (r, b) in enumerate(((0, 2), (0, 4), (0, 4), (0, 1))):
            lib().tl_set_tuning(k, r, b)
        base.close()


def test_bm5_4000_full_deck_matches_reference_goldens():
    """Benchmark 5 (4000x4000, 10 steps): the reference's own golden (tea.problems) and the thesis'
    kernel-call count (42297 calls of cg_calc_w / cg_calc_ur, thesis p.26)."""
    s, summary, hist, u = run_gpu("tea_4000_cg.in")
    from exploringsycl_b200 import get_checking_value
    exp = get_checking_value(os.path.join(GOLDEN, "tea_problems.txt"), s)
    assert abs(100.0 * (summary["temp"] / exp) - 100.0) < 0.001  # field_summary_driver.c:42-43
    assert rel(summary["temp"], exp) < 2e-11
    calls = sum(h["total_iters"] for h in hist)
    assert abs(calls - 42297) <= 10, calls  # +-1 per step from the reduction order
    assert rel(summary["vol"], 100.0) < 1e-12 and rel(summary["mass"], 8401.6) < 1e-12
    print("bm5: calls", calls, "temp", repr(summary["temp"]), "iters", [h["iters_a"] for h in hist])


@pytest.mark.parametrize("name,sid", [("cheby", O.CHEBY), ("ppcg", O.PPCG)])
def test_fused_cheby_ppcg_iterations_are_bit_identical(name, sid):
    """One-pass Chebyshev (iterate + calc_u) and PPCG (calc_ur + calc_sd) kernels == the two-kernel
    sequences of kernel_interface.cpp:258-271 / :314-328, bit for bit."""
    from exploringsycl_b200 import TeaLeaf, read_config
    res = []
    for fused in (0, 1):
        s, states = read_config(os.path.join(DECKS, "tea_250_%s.in" % name))
        s.end_step = 3
        s.fuse_p_into_w = fused
        app = TeaLeaf(s, states)
        summary = app.diffuse()
        res.append((summary, [(h["iters_a"], h["iters_b"], h["error"]) for h in app.history],
                    {f: app.chunk.read(f) for f in (3, 4, 7, 2, 5)}))
        app.close()
    assert res[0][1] == res[1][1]
    assert res[0][0] == res[1][0]
    for f in res[0][2]:
        assert np.array_equal(res[0][2][f][2:-2, 2:-2], res[1][2][f][2:-2, 2:-2]), f
    # u's halo is materialised at the end of the fused phase exactly as the per-iteration updates leave it
    assert np.array_equal(res[0][2][3][1:-1, 1:-1], res[1][2][3][1:-1, 1:-1])
