"""Multi-GPU path (needs >= 2 GPUs; run with `gpurun --gpus 2`): one process per GPU, halo payloads
over NVLink peer stores, scalars through the rank-ordered shared-memory reduction.

  * halo exchange is bit-exact, corners included (L/R completes before B/T, remote_halo_driver.c:24-126)
  * a decomposed CG / Chebyshev / PPCG run matches the oracle's N-chunk run (counts +-1, summary 1e-10)
"""
import os
import socket

import numpy as np
import pytest
import torch.multiprocessing as mp

from oracle import oracle as O
from tl_testutil import DECKS, rel

pytestmark = pytest.mark.gpu


def ngpus():
    try:
        from exploringsycl_b200 import lib
        return lib().tl_device_count()
    except Exception:
        return 0


def free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def halo_worker(rank, world, session, grid, out):
    from exploringsycl_b200 import Chunk, Comms, decompose_field
    from exploringsycl_b200._lib import check, lib
    gx, gy = grid
    comms = Comms(session, rank, world, device=rank)
    d = decompose_field(gx, gy, world, rank)
    ch = Chunk(d["nx"], d["ny"], 2, 10, d["neighbours"], d["left"], d["bottom"], device=rank)
    check(lib().tl_comms_attach_chunk(comms.handle, ch.handle))
    hd = 2
    ok = True
    for depth in (2, 1):
        for rep in range(3):  # host-synchronised between exchanges; the back-to-back case is test_halo_exchange_stress
            jj, kk = np.meshgrid(np.arange(ch.y), np.arange(ch.x), indexing="ij")
            gxx, gyy = d["left"] + kk - hd, d["bottom"] + jj - hd
            truth = (gxx + 10000.0 * gyy + 0.25 * rep).astype(np.float64)
            for fid, scale in ((3, 1.0), (0, -3.0)):
                img = np.full((ch.y, ch.x), -777.0)
                img[hd:-hd, hd:-hd] = scale * truth[hd:-hd, hd:-hd]
                ch.write(fid, img)
            flags = [True, False, False, True, False, False]
            ch.halo_update(comms, flags, depth)
            for fid, scale in ((3, 1.0), (0, -3.0)):
                got = ch.read(fid)
                # expected: reflect the GLOBAL field at external faces, neighbours' data elsewhere
                ex = np.clip(gxx, 0, None)
                rx_ = np.where(gxx < 0, -gxx - 1, np.where(gxx >= gx, 2 * gx - gxx - 1, gxx))
                ry_ = np.where(gyy < 0, -gyy - 1, np.where(gyy >= gy, 2 * gy - gyy - 1, gyy))
                exp = scale * (rx_ + 10000.0 * ry_ + 0.25 * rep)
                lo, hi = hd - depth, -(hd - depth) if hd - depth else None
                sl = (slice(lo, hi), slice(lo, hi))
                ok &= bool(np.array_equal(got[sl], exp[sl]))
    out.put((rank, ok))
    comms.barrier()
    ch.close()
    comms.finalise()


def stress_worker(rank, world, session, grid, reps, skew, out):
    """Back-to-back exchanges without host synchronisation, one rank's host delayed between send and unpack."""
    import ctypes as C
    if skew:
        os.environ["TL_TEST_SKEW"] = skew
    from exploringsycl_b200 import Chunk, Comms, decompose_field
    from exploringsycl_b200._lib import check, lib
    gx, gy = grid
    comms = Comms(session, rank, world, device=rank)
    d = decompose_field(gx, gy, world, rank)
    ch = Chunk(d["nx"], d["ny"], 2, 10, d["neighbours"], d["left"], d["bottom"], device=rank)
    check(lib().tl_comms_attach_chunk(comms.handle, ch.handle))
    bad = C.c_long(-1)
    check(lib().tl_halo_stress(ch.handle, comms.handle, gx, gy, reps, 2, C.byref(bad)))
    out.put((rank, bad.value))
    comms.barrier()
    ch.close()
    comms.finalise()


def deck_worker(rank, world, session, deck_file, over, out, skew=None):
    if skew:
        os.environ["TL_TEST_SKEW"] = skew
    from exploringsycl_b200 import Comms, Settings, TeaLeaf, read_config
    comms = Comms(session, rank, world, device=rank)
    over = dict(over)
    # a mesh override goes in BEFORE the deck is parsed (dx, dy and the states' shrunk extents derive from it)
    preset = Settings(grid_x_cells=over.pop("grid_x_cells", 10), grid_y_cells=over.pop("grid_y_cells", 10))
    s, states = read_config(os.path.join(DECKS, deck_file), preset)
    for k, v in over.items():
        setattr(s, k, v)
    app = TeaLeaf(s, states, comms, device=rank)
    summary = app.diffuse()
    out.put((rank, summary, app.history))
    comms.barrier()
    app.close()
    comms.finalise()


def launch(target, world, args):
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    session = "pytest_gpu_%d_%d" % (os.getpid(), free_port())
    extra = ()
    if target is deck_worker and len(args) == 3:  # (deck, overrides, skew)
        args, extra = args[:2], (args[2],)
    procs = [ctx.Process(target=target, args=(r, world, session) + args + (out,) + extra) for r in range(world)]
    for p in procs:
        p.start()
    res = [out.get(timeout=600) for _ in range(world)]
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    return sorted(res, key=lambda t: t[0])


@pytest.mark.parametrize("world", [2, 4, 8])
@pytest.mark.parametrize("grid", [(64, 48), (101, 67)])
def test_halo_exchange_bit_exact(world, grid):
    if ngpus() < world:
        pytest.skip("needs %d GPUs" % world)
    for rank, ok in launch(halo_worker, world, (grid,)):
        assert ok, "rank %d halo mismatch" % rank


@pytest.mark.parametrize("world", [2, 4, 8])
@pytest.mark.parametrize("skew", [None, "0:300", "1:300"])
def test_halo_exchange_stress(world, skew):
    """200 exchanges (changing field sets, depths and contents, generated and verified on the device) with no host
    synchronisation in between; with `skew` one rank's host sleeps 300 us between the send and the unpack launches
    of every phase, so its neighbours run a full exchange ahead: a receive buffer reused before it was unpacked
    (round-1 race: constant buffer parity per face, no ack) fails this test."""
    if ngpus() < world:
        pytest.skip("needs %d GPUs" % world)
    for rank, bad in launch(stress_worker, world, ((96, 80), 200, skew)):
        assert bad == 0, "rank %d: %d wrong halo cells" % (rank, bad)


@pytest.mark.parametrize("world", [2, 4])
@pytest.mark.parametrize("deck", ["tea_250_cheby.in", "tea_250_ppcg.in"])
def test_cheby_ppcg_multi_rank_are_skew_invariant(world, deck):
    """Chebyshev iterates up to 10 times and PPCG 10 inner steps with a halo exchange and no reduction in between:
    a delayed rank must not change a single bit (advisor finding, round 1)."""
    if ngpus() < world:
        pytest.skip("needs %d GPUs" % world)
    a = launch(deck_worker, world, (deck, {"end_step": 2}))
    b = launch(deck_worker, world, (deck, {"end_step": 2}, "1:200"))
    for (ra, sa, ha), (rb, sb, hb) in zip(a, b):
        assert sa == sb
        assert [(h["iters_a"], h["iters_b"], h["error"]) for h in ha] == \
               [(h["iters_a"], h["iters_b"], h["error"]) for h in hb]


@pytest.mark.parametrize("world", [2, 4, 8])
@pytest.mark.parametrize("deck,solver,fused", [("tea_250_cg.in", O.CG, 2), ("tea_250_cg.in", O.CG, 0),
                                               ("tea_250_cheby.in", O.CHEBY, 1),
                                               ("tea_250_ppcg.in", O.PPCG, 1)])
def test_decomposed_deck_matches_oracle(world, deck, solver, fused):
    """fused=2: two-kernel iteration, r's halo travels; fused=0: three kernels, p's halo travels; 1: auto."""
    if ngpus() < world:
        pytest.skip("needs %d GPUs" % world)
    res = launch(deck_worker, world, (deck, {"end_step": 2, "fuse_p_into_w": fused}))
    ores = O.run_deck(O.make_deck(250, solver=solver, end_step=2, num_chunks=world))
    for rank, summary, hist in res:
        a = [h["iters_a"] for h in hist]
        b = [h["iters_b"] for h in hist]
        assert all(abs(x - y) <= 1 for x, y in zip(a, ores["iters_a"])), (a, ores["iters_a"])
        step = 10 if solver == O.CHEBY else 1
        assert all(abs(x - y) <= step for x, y in zip(b, ores["iters_b"])), (b, ores["iters_b"])
        tol = 1e-10 if (solver == O.CG or b == ores["iters_b"]) else 5e-9
        for k in ("vol", "mass", "ie"):
            assert rel(summary[k], ores[k]) < 1e-10
        assert rel(summary["temp"], ores["temp"]) < tol
    assert all(r[1] == res[0][1] for r in res)  # every rank holds the identical (rank-ordered) sums


@pytest.mark.parametrize("world", [2, 4, 8])
def test_fused_and_three_kernel_multi_rank_loops_are_bit_identical(world):
    if ngpus() < world:
        pytest.skip("needs %d GPUs" % world)
    a = launch(deck_worker, world, ("tea_250_cg.in", {"end_step": 2, "fuse_p_into_w": 2}))
    b = launch(deck_worker, world, ("tea_250_cg.in", {"end_step": 2, "fuse_p_into_w": 0}))
    for (ra, sa, ha), (rb, sb, hb) in zip(a, b):
        assert sa == sb
        assert [h["iters_a"] for h in ha] == [h["iters_a"] for h in hb]
        assert [h["error"] for h in ha] == [h["error"] for h in hb]


@pytest.mark.parametrize("binary", ["tealeaf_cuda", "tealeaf_cuda_resident"])
def test_reference_host_on_two_ranks(binary, tmp_path):
    """The unmodified reference host, one process per GPU (launched with RANK / WORLD_SIZE / LOCAL_RANK /
    MASTER_PORT in the environment), with comms_b200.cpp in place of the MPI comms.c: decomposed CG deck
    vs the oracle's 2-chunk run.  tealeaf_cuda = plugin path (host buffers through the mailboxes),
    tealeaf_cuda_resident = DIFFUSE_OVERLOAD (device-resident loop, NVLink peer stores)."""
    import re
    import shutil
    import subprocess
    import sys
    from tl_testutil import GOLDEN, ROOT
    if ngpus() < 2:
        pytest.skip("needs 2 GPUs")
    exe = os.path.join(ROOT, "c_kernels", "cuda", "bin", binary)
    if not os.path.exists(exe):
        pytest.skip("%s not built" % binary)
    shutil.copy(os.path.join(DECKS, "tea_250_cg.in"), tmp_path / "tea.in")
    shutil.copy(os.path.join(GOLDEN, "tea_problems.txt"), tmp_path / "tea.problems")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--no-python", "--nnodes=1", "--nproc-per-node", "2",
           "--master-addr", "127.0.0.1", "--master-port", str(free_port()), exe]
    r = subprocess.run(cmd, cwd=tmp_path, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "PASSED" in r.stdout
    cg = [int(v) for v in re.findall(r"(?m)^CG:\s+(\d+) iterations", r.stdout)]
    actual = float(re.search(r"Actual\s+(\S+)", r.stdout).group(1))
    ores = O.run_deck(O.make_deck(250, num_chunks=2))
    assert len(cg) == 10 and all(abs(a - b) <= 1 for a, b in zip(cg, ores["iters_a"])), (cg, ores["iters_a"])
    assert rel(actual, ores["temp"]) < 1e-10


@pytest.mark.parametrize("world", [2, 4, 8])
@pytest.mark.parametrize("fused", [2, 0])
def test_decomposed_cg_is_bit_identical_to_the_oracle_n_chunk_run(world, fused):
    """With the oracle replaying the GPU summation tree per chunk and adding the chunk sums in rank order
    (exactly what the NVLink slot reduction does), a decomposed solve matches in every bit."""
    if ngpus() < world:
        pytest.skip("needs %d GPUs" % world)
    res = launch(deck_worker, world, ("tea_250_cg.in", {"end_step": 3, "fuse_p_into_w": fused}))
    ores = O.run_deck(O.make_deck(250, end_step=3, num_chunks=world), gpu_sum_order=True)
    for rank, summary, hist in res:
        assert [h["iters_a"] for h in hist] == ores["iters_a"]
        assert [h["error"] for h in hist] == ores["error"]
        assert [summary[k] for k in ("vol", "mass", "ie", "temp")] == [ores[k] for k in ("vol", "mass", "ie", "temp")]


@pytest.mark.parametrize("world", [2, 4, 8])
def test_wide_chunks_with_unaligned_tile_groups_are_bit_identical_to_the_oracle(world):
    """1200 x 600 cells: every chunk is several column tiles wide (3 at world 4), so the 64-tile groups of the grid
    reduction straddle tile rows and the group-wise forwarding of the parked halo columns (ForwardColumns,
    tl_device.cuh) has to work out which rows of the left / right edge tiles each group owns.  Fused and three-kernel
    loops against the oracle's N-chunk run, bit for bit."""
    if ngpus() < world:
        pytest.skip("needs %d GPUs" % world)
    over = {"end_step": 2, "grid_x_cells": 1200, "grid_y_cells": 600}
    ores = O.run_deck(O.make_deck(1200, 600, end_step=2, num_chunks=world), gpu_sum_order=True)
    for fused in (2, 0):
        res = launch(deck_worker, world, ("tea_250_cg.in", dict(over, fuse_p_into_w=fused)))
        for rank, summary, hist in res:
            assert [h["iters_a"] for h in hist] == ores["iters_a"], (fused, rank)
            assert [h["error"] for h in hist] == ores["error"], (fused, rank)
            assert [summary[k] for k in ("vol", "mass", "ie", "temp")] == [ores[k] for k in ("vol", "mass", "ie", "temp")]
