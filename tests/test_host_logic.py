"""Host logic that runs without a GPU: deck parser, decomposition (bit-exact vs the oracle's
restatement of initialise.c:34-134), checking-value lookup."""
import ctypes as C
import os

import numpy as np
import pytest

from oracle import oracle as O
from tl_testutil import DECKS, GOLDEN


def test_read_config_standard_deck():
    from exploringsycl_b200 import read_config
    s, st = read_config(os.path.join(DECKS, "tea_250_cg.in"))
    assert (s.grid_x_cells, s.grid_y_cells, s.end_step, s.max_iters) == (250, 250, 10, 10000)
    assert s.dt_init == 0.004 and s.eps == 1.0e-15 and s.solver == 1
    assert s.grid_x_max == 10.0 and s.dx == 10.0 / 250
    assert len(st) == 5 and st[0].density == 100.0 and st[0].energy == 0.0001
    # extents shrunk by dx/100 (parse_config.c:253-260)
    assert st[1].x_min == 0.0 + s.dx / 100.0 and st[1].x_max == 1.0 - s.dx / 100.0
    assert st[4].y_min == 7.0 + s.dy / 100.0 and st[4].y_max == 8.0 - s.dy / 100.0
    s2, _ = read_config(os.path.join(DECKS, "tea_250_ppcg.in"))
    assert s2.solver == 3
    s3, _ = read_config(os.path.join(DECKS, "tea_10_jacobi.in"))
    assert s3.solver == 0 and s3.grid_x_cells == 10


def test_read_config_quirks(tmp_path):
    from exploringsycl_b200 import read_config
    p = tmp_path / "tea.in"
    p.write_text("*tea\nstate 1 density=1.0 energy=2.0\n  x_cells=32\ny_cells=48\nepslim=1e-4\neps 1e-12\n"
                 "use_chebyshev\nerrswitch\npresteps=12\nppcg_inner_steps=7\ncoefficient_inverse_density\n"
                 "tl_something 5\nprofiler_on\n*endtea\n")
    s, st = read_config(str(p))
    assert s.grid_x_cells == 32 and s.grid_y_cells == 48
    assert s.eps_lim == 1e-4 and s.eps == 1e-12  # "epslim" is tested before "eps" (parse_config.c:114-119)
    assert s.solver == 2 and s.error_switch and s.presteps == 12 and s.ppcg_inner_steps == 7
    assert s.coefficient == 2 and len(st) == 1


@pytest.mark.parametrize("grid", [(4000, 4000), (250, 250), (101, 67), (8000, 16000), (16000, 16000)])
@pytest.mark.parametrize("n", [1, 2, 3, 4, 6, 8])
def test_decomposition_bit_exact(grid, n):
    from exploringsycl_b200 import decompose_field
    gx, gy = grid
    xc, yc = C.c_int(), C.c_int()
    z = [np.zeros(n, dtype=np.int32) for _ in range(4)]
    nb = np.zeros(4 * n, dtype=np.int32)
    assert O.lib().orc_decompose(gx, gy, n, C.byref(xc), C.byref(yc), *z, nb) == 0
    left, right, bottom, top = z
    for c in range(n):
        d = decompose_field(gx, gy, n, c)
        assert (d["x_chunks"], d["y_chunks"]) == (xc.value, yc.value)
        assert d["left"] == left[c] and d["bottom"] == bottom[c]
        assert d["nx"] == right[c] - left[c] and d["ny"] == top[c] - bottom[c]
        assert d["neighbours"] == list(nb[4 * c:4 * c + 4])


def test_named_decompositions():
    from exploringsycl_b200 import decompose_field
    # N=2 -> 1x2, N=4 -> 2x2, N=8 -> 2x4 on square meshes
    assert [decompose_field(4000, 4000, n, 0)[k] for n in (2, 4, 8) for k in ("x_chunks", "y_chunks")] == \
        [1, 2, 2, 2, 2, 4]
    d = decompose_field(4000, 4000, 8, 3)
    assert (d["nx"], d["ny"]) == (2000, 1000) and d["neighbours"] == [2, -1, 1, 5]


def test_checking_value():
    from exploringsycl_b200 import Settings, get_checking_value
    s = Settings(grid_x_cells=4000, grid_y_cells=4000, end_step=10)
    assert get_checking_value(os.path.join(GOLDEN, "tea_problems.txt"), s) == 9.5462351582214282e+01
    s.end_step = 3
    assert get_checking_value(os.path.join(GOLDEN, "tea_problems.txt"), s) is None


QUIRK_DECKS = {
    "quirks": "*tea\nstate 1 density=1.0 energy=2.0\nstate 2 density=0.5 energy=3.5 geometry=rectangle xmin=1.0 xmax=2.0 "
              "ymin=0.5 ymax=3.0\n  x_cells=32\ny_cells=48\nxmin=-1.0\nxmax=7.0\nymax=3.0\nepslim=1e-4\neps 1e-12\n"
              "use_chebyshev\nerrswitch\npresteps=12\nppcg_inner_steps=7\ncoefficient_inverse_density\n"
              "initial_timestep=0.01\nend_step=3\nmax_iters=777\nsummary_frequency=2\ntl_something 5\nprofiler_on\n"
              "*endtea\n",
    "circle": "*tea\nstate 1 density=100.0 energy=0.0001\nstate 2 density=0.1 energy=25.0 geometry=circular xmin=5.0 "
              "xmax=0.0 ymin=5.0 ymax=0.0 radius=2.5\nstate 3 density=7 energy=1 geometry=point xmin=1.0 xmax=0 ymin=2.0 "
              "ymax=0\nx_cells=20\ny_cells=30\nxmax=10.0\nymax=10.0\nend_step=1\nuse_jacobi\n*endtea\n",
}


@pytest.mark.skipif(not os.path.exists(os.path.join(os.path.dirname(O.REF_BIN), "print_config")),
                    reason="oracle/_ref/print_config not built (needs /root/reference at build time)")
@pytest.mark.parametrize("deck", ["tea_250_cg.in", "tea_250_cheby.in", "tea_250_ppcg.in", "tea_10_jacobi.in",
                                  "tea_4000_cg.in", "quirks", "circle"])
def test_read_config_matches_the_reference_parser(deck, tmp_path):
    """exploringsycl_b200.read_config vs the reference's own parse_config.c (compiled in place into
    oracle/_ref/print_config): every Settings field and every State, bit for bit."""
    import subprocess
    from exploringsycl_b200 import read_config
    path = tmp_path / "tea.in"
    if deck in QUIRK_DECKS:
        path.write_text(QUIRK_DECKS[deck])
    else:
        path.write_text(open(os.path.join(DECKS, deck)).read())
    out = subprocess.run([os.path.join(os.path.dirname(O.REF_BIN), "print_config")], cwd=tmp_path,
                         capture_output=True, text=True, timeout=60).stdout
    ref, ref_states = {}, []
    for line in out.splitlines():
        tok = line.split()
        if tok[0] == "state":
            ref_states.append([float(t) for t in tok[2:]])
        else:
            ref[tok[0]] = float(tok[1])
    s, st = read_config(str(path))
    for key in ("grid_x_cells", "grid_y_cells", "end_step", "max_iters", "presteps", "ppcg_inner_steps",
                "summary_frequency", "halo_depth", "num_states", "solver", "coefficient", "dt_init", "eps",
                "eps_lim", "end_time", "grid_x_min", "grid_y_min", "grid_x_max", "grid_y_max", "dx", "dy"):
        assert float(getattr(s, key)) == ref[key], key
    assert float(s.error_switch) == ref["error_switch"] and float(s.check_result) == ref["check_result"]
    assert len(st) == len(ref_states)
    for n, (a, b) in enumerate(zip(st, ref_states)):
        mine = [a.density, a.energy] if n == 0 else [a.density, a.energy, float(a.geometry), a.x_min, a.y_min,
                                                      a.x_max, a.y_max]
        assert mine == b, (n, mine, b)


def test_decomposition_invariants_property():
    """Hypothesis: for any mesh and rank count the chunks tile the mesh exactly once, neighbours are mutual,
    and the library agrees with the oracle's restatement of initialise.c:34-134."""
    from hypothesis import given, settings as hsettings, strategies as st
    from exploringsycl_b200 import decompose_field

    @hsettings(max_examples=60, deadline=None)
    @given(st.integers(8, 3000), st.integers(8, 3000), st.sampled_from([1, 2, 3, 4, 5, 6, 7, 8]))
    def check(gx, gy, n):
        ds = [decompose_field(gx, gy, n, c) for c in range(n)]
        cover = np.zeros((gy, gx), dtype=np.int32)
        for d in ds:
            cover[d["bottom"]:d["bottom"] + d["ny"], d["left"]:d["left"] + d["nx"]] += 1
        assert cover.min() == 1 and cover.max() == 1
        opp = {0: 1, 1: 0, 2: 3, 3: 2}
        for c, d in enumerate(ds):
            for face, nb in enumerate(d["neighbours"]):
                if nb != -1:
                    assert ds[nb]["neighbours"][opp[face]] == c
                    if face in (0, 1):  # left/right neighbours share the row extent
                        assert ds[nb]["bottom"] == d["bottom"] and ds[nb]["ny"] == d["ny"]
                    else:
                        assert ds[nb]["left"] == d["left"] and ds[nb]["nx"] == d["nx"]
        assert all(d["x_chunks"] * d["y_chunks"] == n for d in ds)

    check()


# ------------------------------------------------------------------------------------------------
# command-line and deck hygiene (opt-in; the default reader stays pinned to the reference's quirks)
# ------------------------------------------------------------------------------------------------
QUIRKY_DECK = """*tea
state 1 density=100.0 energy=0.0001
state 2 density=0.1 energy=25.0 geometry=circular xmin=-2.5 ymin=1.0 radius=3.0
x_cells=10
y_cells=40
xmin=-5.0
ymin=0.0
xmax=5.0
ymax=10.0
initial_timestep=0.004
end_step=3
tl_max_iters=777
tl_use_ppcg
tl_ppcg_inner_steps=4
eps 1.0e-12
epslim 1.0e-3
profiler_on
frobnicate=1
*endtea
"""


def test_clean_deck_reader_versus_reference_quirks(tmp_path):
    from exploringsycl_b200 import Settings, TeaLeafError, read_config
    from exploringsycl_b200.tealeaf import CIRCULAR, PPCG_SOLVER
    deck = tmp_path / "tea.in"
    deck.write_text(QUIRKY_DECK)
    # the reference's reader (parse_config.c): sign dropped, tl_ keys and profiler_on ignored, circular state without
    # xmax is an error ("Failed to find a value")
    with pytest.raises(TeaLeafError):
        read_config(str(deck))
    deck.write_text(QUIRKY_DECK.replace("radius=3.0", "radius=3.0 xmax=0.0 ymax=0.0"))
    q, _ = read_config(str(deck))
    assert q.grid_x_min == 5.0 and q.max_iters == 10000 and q.solver == 1 and q.ppcg_inner_steps == 10
    assert q.grid_x_cells == 10 and q.grid_y_cells == 40
    # the clean reader
    deck.write_text(QUIRKY_DECK)
    s, states = read_config(str(deck), hygiene=True)
    assert s.grid_x_min == -5.0 and s.grid_x_max == 5.0 and s.dx == 1.0
    assert s.max_iters == 777 and s.solver == PPCG_SOLVER and s.ppcg_inner_steps == 4
    assert s.eps == 1.0e-12 and s.eps_lim == 1.0e-3 and s.profiler_on
    assert s.deck_warnings == ["unknown key 'frobnicate'"]
    assert states[1].geometry == CIRCULAR and states[1].radius == 3.0
    assert states[1].x_min == -2.5 + s.dx / 100.0 and states[1].y_min == 1.0 + s.dy / 100.0
    # x_cells = 10 in the deck is honoured even though 10 is the default, and the command line wins over the deck
    s2, _ = read_config(str(deck), Settings(grid_x_cells=64), hygiene=True)
    assert s2.grid_x_cells == 64 and s2.grid_y_cells == 40
    for bad in ("state 2 density=0.1", "x_cells=0"):
        deck.write_text(QUIRKY_DECK.replace("x_cells=10", bad))
        with pytest.raises(TeaLeafError):
            read_config(str(deck), hygiene=True)


def test_settings_overload_quirk_and_fix():
    """main.c:76-85: `-x 4000` calls atoi("-x") -> 0 cells.  The hygienic overload reads the value."""
    from exploringsycl_b200 import Settings, TeaLeafError, settings_overload
    from exploringsycl_b200.tealeaf import CHEBY_SOLVER
    s = settings_overload(Settings(), ["tealeaf", "-x", "4000", "-y", "2000", "-s", "cheby"])
    assert (s.grid_x_cells, s.grid_y_cells, s.solver) == (0, 0, CHEBY_SOLVER)
    s = settings_overload(Settings(), ["tealeaf", "-x", "4000", "-y", "2000", "--solver", "cheby"], hygiene=True)
    assert (s.grid_x_cells, s.grid_y_cells, s.solver) == (4000, 2000, CHEBY_SOLVER)
    with pytest.raises(TeaLeafError):
        settings_overload(Settings(), ["tealeaf", "-x", "abc"], hygiene=True)
    with pytest.raises(TeaLeafError):
        settings_overload(Settings(), ["tealeaf", "-s", "multigrid"], hygiene=True)
    assert settings_overload(Settings(), ["tealeaf", "-x"]).grid_x_cells == 10  # dangling option: ignored (main.c:77)


def test_write_to_visit(tmp_path):
    """shared.c:114-150: header lines and the raw doubles."""
    import numpy as np
    from exploringsycl_b200 import write_to_visit
    data = np.arange(12.0).reshape(3, 4)
    bov = write_to_visit(4, 3, 8, 16, data, "density", 7, 0.028, directory=str(tmp_path))
    txt = open(bov).read().splitlines()
    assert txt[0] == "TIME: 0.0280" and txt[1] == "DATA_FILE: density7.dat" and txt[2] == "DATA_SIZE: 4 3 1"
    assert "DATA_FORMAT: DOUBLE" in txt and "CENTERING: zone" in txt and "BRICK_ORIGIN: 8. 16. 0." in txt
    assert np.array_equal(np.fromfile(tmp_path / "density7.dat", dtype="<f8").reshape(3, 4), data)
