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
