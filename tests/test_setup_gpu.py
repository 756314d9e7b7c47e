"""Setup kernels and whole decks the round-1 tests did not reach:
  * set_chunk_state with circular and point geometries (set_chunk_state.cpp:30-92) -- bit-exact fields;
  * a whole deck with coefficient_inverse_density (RECIP_CONDUCTIVITY, cg.cpp:34-36, jacobi.cpp:29-31);
  * a deck mixing rectangle / circle / point states through the deck parser;
  * the vectorised streaming kernels on odd shapes and odd halo depths (double2 alignment shifts)."""
import ctypes as C
import os

import numpy as np
import pytest

from oracle import oracle as O
from tl_testutil import DECKS, HD, dbl, rel, rng_fields, upload

pytestmark = pytest.mark.gpu

DECK_MIXED = """*tea
state 1 density=100.0 energy=0.0001
state 2 density=0.1 energy=25.0 geometry=circular xmin=3.0 ymin=4.0 xmax=0.0 ymax=0.0 radius=1.7
state 3 density=0.2 energy=0.1 geometry=rectangle xmin=5.0 xmax=9.0 ymin=6.0 ymax=8.0
state 4 density=7.5 energy=3.0 geometry=point xmin=%s ymin=%s xmax=0.0 ymax=0.0
x_cells=120
y_cells=90
xmin=0.0
ymin=0.0
xmax=10.0
ymax=10.0
initial_timestep=0.004
end_step=3
max_iters=10000
use_cg
%s
eps 1.0e-15
use_c_kernels
*endtea
"""


def oracle_setup(s, states, nx, ny, hd):
    """initialise_application of the oracle kernels for one chunk covering the mesh"""
    x, y = nx + 2 * hd, ny + 2 * hd
    L = O.lib()
    vx, vy, cx, cy = np.zeros(x + 1), np.zeros(y + 1), np.zeros(x), np.zeros(y)
    vol = np.zeros((y, x))
    L.orc_set_chunk_data(x, y, hd, s.grid_x_min, s.grid_y_min, s.dx, s.dy, vx, vy, cx, cy, vol)
    e0, den, u = np.zeros((y, x)), np.zeros((y, x)), np.zeros((y, x))
    L.orc_set_chunk_initial_state(x, y, states[0].energy, states[0].density, e0, den)
    for st in states[1:]:
        L.orc_set_chunk_state(x, y, hd, st.geometry, st.density, st.energy, st.x_min, st.y_min, st.x_max, st.y_max,
                              st.radius, e0, den, u, cx, cy, vx, vy)
    return e0, den, u, vol


@pytest.mark.parametrize("coef_line", ["", "coefficient_inverse_density"])
def test_circular_and_point_states_and_recip_deck(tmp_path, coef_line):
    from exploringsycl_b200 import Chunk, TeaLeaf, read_config
    from exploringsycl_b200.tealeaf import CIRCULAR, POINT
    nx, ny = 120, 90
    dx, dy = 10.0 / nx, 10.0 / ny
    # a point state applies where a VERTEX equals the (shifted) position exactly (set_chunk_state.cpp:62-65): put the
    # deck value so that value + dx/100 lands on vertex 37 / 21 of the chunk
    def deck_value(target, shift):  # v with float(v) + shift == target exactly
        v = target - shift
        for _ in range(64):
            if v + shift == target:
                return v
            v = np.nextafter(v, np.inf if v + shift < target else -np.inf)
        raise AssertionError("no representable deck value")
    px, py = deck_value(0.0 + dx * 37.0, dx / 100.0), deck_value(0.0 + dy * 21.0, dy / 100.0)
    deck = tmp_path / "tea.in"
    deck.write_text(DECK_MIXED % (repr(px), repr(py), coef_line))
    s, states = read_config(str(deck))
    assert [st.geometry for st in states[1:]] == [CIRCULAR, 0, POINT]
    assert s.coefficient == (2 if coef_line else 1)
    # --- kernel level: fields after initialise_application, bit for bit ---
    ch = Chunk(nx, ny, HD, 10)
    ch.run_set_chunk_data(s)
    ch.run_set_chunk_state(s, states)
    e0, den, u, vol = oracle_setup(s, states, nx, ny, HD)
    assert np.array_equal(ch.read(1), e0) and np.array_equal(ch.read(0), den) and np.array_equal(ch.read(11), vol)
    interior1 = (slice(1, -1), slice(1, -1))
    assert np.array_equal(ch.read(3)[interior1], u[interior1])
    n_circle = int((den == 0.1).sum())
    assert n_circle > 50, "the circular state must cover cells"
    assert int((den == 7.5).sum()) == 1, "the point state must hit exactly one cell"
    ch.close()
    # --- whole deck vs the oracle (same states, same coefficient) ---
    od = O.make_deck(nx, ny, end_step=3, states=[(100.0, 0.0001, None)], coefficient=s.coefficient)
    od.num_states = len(states)
    for n, st in enumerate(states):
        q = od.states[n]
        q.geometry, q.density, q.energy, q.radius = st.geometry, st.density, st.energy, st.radius
        if n:  # the oracle driver applies the dx/100 shrink itself (parse_config.c:253-260)
            q.x_min, q.y_min = st.x_min - s.dx / 100.0, st.y_min - s.dy / 100.0
            q.x_max, q.y_max = st.x_max + s.dx / 100.0, st.y_max + s.dy / 100.0
    ores = O.run_deck(od)
    app = TeaLeaf(s, states)
    summary = app.diffuse()
    hist = app.history
    app.close()
    assert all(abs(a["iters_a"] - b) <= 1 for a, b in zip(hist, ores["iters_a"])), ([h["iters_a"] for h in hist], ores["iters_a"])
    for k in ("vol", "mass", "ie", "temp"):
        assert rel(summary[k], ores[k]) < 1e-10, k


@pytest.mark.parametrize("nx,ny,hd", [(5, 3, 1), (37, 23, 1), (64, 17, 3), (513, 300, 2), (1000, 77, 2)])
def test_vectorised_streaming_kernels_on_odd_shapes(nx, ny, hd):
    """copy (all cells / interior), finalise, 2-norm, field summary, jacobi: the two-columns-per-thread skeleton must
    honour every range start and row end, also when off + k_lo is odd (halo depth 1 and 3)."""
    from exploringsycl_b200 import Chunk
    x, y = nx + 2 * hd, ny + 2 * hd
    ch = Chunk(nx, ny, hd, 10)
    f = rng_fields(nx, ny, seed=7 * nx + ny, hd=hd)
    upload(ch, f)
    L = O.lib()
    ch.run_store_energy()  # energy = energy0, ALL cells
    assert np.array_equal(ch.read(2), f["energy0"])
    ch.run_copy_u()        # u0 = u, interior only
    exp = f["u0"].copy()
    L.orc_copy_u(x, y, hd, f["u"], exp)
    assert np.array_equal(ch.read(6), exp)
    n = ch.run_calculate_2norm(7)
    o = dbl(); L.orc_calculate_2norm(x, y, hd, f["r"], C.byref(o))
    assert rel(n, o.value) < 1e-13
    got = ch.run_field_summary()
    v = [dbl() for _ in range(4)]
    L.orc_field_summary(x, y, hd, f["volume"], f["density"], f["energy0"], f["u"], *[C.byref(q) for q in v])
    assert all(rel(a, b.value) < 1e-13 for a, b in zip(got, v))
    ch.run_finalise()      # energy = u / density, interior
    en = f["energy0"].copy()
    L.orc_finalise(x, y, hd, f["u"], f["density"], en)
    assert np.array_equal(ch.read(2), en)
    err = ch.run_jacobi_iterate()
    u, r = f["u"].copy(), f["r"].copy()
    oe = dbl(); L.orc_jacobi_iterate(x, y, hd, u, exp, r, f["kx"], f["ky"], C.byref(oe))  # u0 as copy_u left it
    assert np.array_equal(ch.read(3), u) and np.array_equal(ch.read(7), r)
    assert rel(err, oe.value) < 1e-13
    ch.close()
