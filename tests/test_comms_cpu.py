"""N > 1 host path on CPU (world_size 2 and 4, gloo for process-group plumbing).

The product's comms layer (tl_comms_*: shared-memory rendezvous, barrier, rank-ordered sum / min,
send_recv mailboxes) and decomposition run for real in `host_only` mode; the per-chunk compute is
stood in for by the oracle's kernels (tests may use the oracle; the product never does).  Each rank
drives its chunk through the reference's cg_driver / remote_halo_driver call order and the result
must equal the oracle's in-process N-chunk run bit for bit (same kernels, same rank-ordered sums).
"""
import ctypes as C
import os
import socket

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import oracle as O

HD = 2


def free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


class OracleChunk:
    """One rank's chunk computed by the oracle kernels (test stand-in for the GPU chunk)."""

    def __init__(self, deck, d, comms):
        self.comms, self.deck, self.d = comms, deck, d
        self.x, self.y = d["nx"] + 2 * HD, d["ny"] + 2 * HD
        self.nb = d["neighbours"]
        n = (self.y, self.x)
        self.f = {k: np.zeros(n) for k in ("density", "energy0", "energy", "u", "u0", "p", "r", "w", "kx", "ky",
                                           "sd", "volume")}
        self.fields = [False] * 6
        L = O.lib()
        dx = (deck.xmax - deck.xmin) / deck.x_cells
        dy = (deck.ymax - deck.ymin) / deck.y_cells
        self.dx, self.dy = dx, dy
        vx, vy = np.zeros(self.x + 1), np.zeros(self.y + 1)
        cx, cy = np.zeros(self.x), np.zeros(self.y)
        L.orc_set_chunk_data(self.x, self.y, HD, deck.xmin + dx * d["left"], deck.ymin + dy * d["bottom"], dx, dy,
                             vx, vy, cx, cy, self.f["volume"])
        L.orc_set_chunk_initial_state(self.x, self.y, deck.states[0].energy, deck.states[0].density,
                                      self.f["energy0"], self.f["density"])
        for i in range(1, deck.num_states):
            s = deck.states[i]
            L.orc_set_chunk_state(self.x, self.y, HD, 0, s.density, s.energy, s.x_min + dx / 100, s.y_min + dy / 100,
                                  s.x_max - dx / 100, s.y_max - dy / 100, 0.0, self.f["energy0"],
                                  self.f["density"], self.f["u"], cx, cy, vx, vy)

    ORDER = ["density", "energy0", "energy", "u", "p", "sd"]

    def halo_update(self, depth):  # halo_update_driver.c + remote_halo_driver.c over the product comms
        if not any(self.fields):
            return
        L = O.lib()
        names = [n for n, f in zip(self.ORDER, self.fields) if f]
        for faces in ((O.LEFT, O.RIGHT), (O.BOTTOM, O.TOP)):
            tags = {O.LEFT: (0, 1), O.RIGHT: (1, 0), O.BOTTOM: (0, 1), O.TOP: (1, 0)}
            bufs = {}
            for face in faces:  # pack + post (send_recv_message is post+receive in one call here)
                if self.nb[face] == O.EXTERNAL:
                    continue
                per = depth * (self.y if face <= O.RIGHT else self.x)
                send = np.zeros(per * len(names))
                for k, n in enumerate(names):
                    b = np.zeros(per)
                    L.orc_pack(self.x, self.y, HD, depth, face, self.f[n], b)
                    send[k * per:(k + 1) * per] = b
                bufs[face] = (send, np.zeros_like(send), per)
            for face in faces:  # send_recv_message posts (remote_halo_driver.c:32-49) ...
                if face in bufs:
                    self.comms.post(bufs[face][0], self.nb[face], tags[face][0])
            for face in faces:  # ... wait_for_requests completes (remote_halo_driver.c:55)
                if face in bufs:
                    self.comms.recv(bufs[face][1], self.nb[face], tags[face][1])
            for face in faces:
                if face in bufs:
                    send, recv, per = bufs[face]
                    for k, n in enumerate(names):
                        L.orc_unpack(self.x, self.y, HD, depth, face, self.f[n],
                                     np.ascontiguousarray(recv[k * per:(k + 1) * per]))
        for n in ("density", "p", "energy0", "energy", "u", "sd"):  # kernel_interface.cpp:120-126
            if self.fields[self.ORDER.index(n)]:
                for face in (O.LEFT, O.RIGHT, O.TOP, O.BOTTOM):
                    if self.nb[face] == O.EXTERNAL:
                        L.orc_local_halo(self.x, self.y, HD, depth, face, self.f[n])

    def set_fields(self, *names):
        self.fields = [n in names for n in self.ORDER]

    def cg_timestep(self, deck):
        """diffuse.c:23-64 + cg_driver.c + solve_finished_driver.c for this rank."""
        L, f, c = O.lib(), self.f, self.comms
        dt = c.min_over_ranks(deck.dt_init)
        rx, ry = dt / (self.dx * self.dx), dt / (self.dy * self.dy)
        self.set_fields("energy", "density")
        self.halo_update(2)
        rro = C.c_double(0.0)
        L.orc_cg_init(self.x, self.y, HD, deck.coefficient, rx, ry, f["density"], f["energy"], f["u"], f["p"],
                      f["r"], f["w"], f["kx"], f["ky"], C.byref(rro))
        self.set_fields("u", "p")
        self.halo_update(1)
        rro = c.sum_over_ranks(rro.value)
        L.orc_copy_u(self.x, self.y, HD, f["u"], f["u0"])
        tt = 0
        for tt in range(deck.max_iters):
            pw = C.c_double(0.0)
            L.orc_cg_calc_w(self.x, self.y, HD, f["p"], f["kx"], f["ky"], f["w"], C.byref(pw))
            alpha = rro / c.sum_over_ranks(pw.value)
            rrn = C.c_double(0.0)
            L.orc_cg_calc_ur(self.x, self.y, HD, alpha, f["p"], f["w"], f["u"], f["r"], C.byref(rrn))
            rrn = c.sum_over_ranks(rrn.value)
            L.orc_cg_calc_p(self.x, self.y, HD, rrn / rro, f["r"], f["p"])
            rro = rrn
            self.halo_update(1)
            if abs(rrn) ** 0.5 < deck.eps:
                break
        L.orc_calculate_residual(self.x, self.y, HD, f["u"], f["u0"], f["kx"], f["ky"], f["r"])
        n = C.c_double(0.0)
        L.orc_calculate_2norm(self.x, self.y, HD, f["r"], C.byref(n))
        c.sum_over_ranks(n.value)
        L.orc_finalise(self.x, self.y, HD, f["u"], f["density"], f["energy"])
        self.fields[self.ORDER.index("energy")] = True
        self.halo_update(1)
        return tt

    def field_summary(self):
        v = [C.c_double() for _ in range(4)]
        O.lib().orc_field_summary(self.x, self.y, HD, self.f["volume"], self.f["density"], self.f["energy0"],
                                  self.f["u"], *[C.byref(q) for q in v])
        return [self.comms.sum_over_ranks(q.value) for q in v]


def worker(rank, world, port, session, n, steps, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from exploringsycl_b200 import Comms, decompose_field
    comms = Comms(session, rank, world, host_only=True)
    res = {}
    # --- collectives ---
    comms.barrier()
    vals = [0.1 * (r + 1) + 1e-17 * r for r in range(world)]
    s = comms.sum_over_ranks(vals[rank])
    exp = vals[0]
    for v in vals[1:]:
        exp += v  # rank order
    res["sum_ok"] = (s == exp)
    res["min_ok"] = (comms.min_over_ranks(5.0 - rank) == 5.0 - (world - 1))
    # --- ring send/recv, twice (mailbox reuse) ---
    ok = True
    if world > 1:
        for rep in range(3):
            right, left = (rank + 1) % world, (rank - 1) % world
            send = np.full(7, 100.0 * rank + rep)
            recv = np.zeros(7)
            if world == 2:
                comms.send_recv_message(send, recv, right, 0, 0)
                ok &= bool(np.all(recv == 100.0 * right + rep))
            else:
                # tag 1 towards the right neighbour, tag 0 towards the left (remote_halo_driver.c tags)
                r2 = np.zeros(7)
                comms.post(send, right, 1)
                comms.post(send, left, 0)
                comms.recv(recv, right, 0)
                comms.recv(r2, left, 1)
                ok &= bool(np.all(recv == 100.0 * right + rep) and np.all(r2 == 100.0 * left + rep))
    res["ring_ok"] = ok
    # --- a CG deck through the product comms + decomposition ---
    deck = O.make_deck(n, end_step=steps, num_chunks=world)
    d = decompose_field(n, n, world, rank)
    ch = OracleChunk(deck, d, comms)
    ch.set_fields("density", "energy0", "energy")
    ch.halo_update(2)
    O.lib().orc_store_energy(ch.x, ch.y, ch.f["energy0"], ch.f["energy"])
    res["iters"] = [ch.cg_timestep(deck) for _ in range(steps)]
    res["summary"] = ch.field_summary()
    gathered = [None] * world
    dist.all_gather_object(gathered, res)
    if rank == 0:
        out.put(gathered)
    comms.barrier()
    comms.finalise()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 4])
def test_cg_over_product_comms_matches_oracle_n_chunk_run(world):
    n, steps = 64, 2
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = free_port()
    session = "pytest_%d_%d" % (os.getpid(), world)
    procs = [ctx.Process(target=worker, args=(r, world, port, session, n, steps, out)) for r in range(world)]
    for p in procs:
        p.start()
    gathered = out.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    ref = O.run_deck(O.make_deck(n, end_step=steps, num_chunks=world))
    for r in gathered:
        assert r["sum_ok"] and r["min_ok"] and r["ring_ok"]
        assert r["iters"] == ref["iters_a"]
        assert r["summary"] == [ref["vol"], ref["mass"], ref["ie"], ref["temp"]]  # bit-exact


def _err_worker(rank, world, session, out):
    from exploringsycl_b200 import Comms, TeaLeafError
    res = {}
    for rep in range(2):  # the same session name can be created again after a clean teardown
        comms = Comms(session, rank, world, host_only=True)
        other = 1 - rank
        send = np.full(5 + rank, float(rank))  # rank 0 sends 5 doubles, rank 1 sends 6
        recv = np.zeros(5 + rank)               # ... and each expects its OWN length: mismatch
        try:
            comms.send_recv_message(send, recv, other, 0, 0)
            res["rep%d" % rep] = "no error"
        except TeaLeafError as e:
            res["rep%d" % rep] = str(e)
        comms.barrier()
        comms.finalise()
    out.put((rank, res))


def test_message_length_mismatch_is_reported_and_session_is_reusable():
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    session = "pytest_err_%d" % os.getpid()
    procs = [ctx.Process(target=_err_worker, args=(r, 2, session, out)) for r in range(2)]
    for p in procs:
        p.start()
    res = dict(out.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank in (0, 1):
        for rep in (0, 1):
            assert "message length mismatch" in res[rank]["rep%d" % rep]


def _abort_worker(rank, world, session, out):
    import time
    from exploringsycl_b200 import Comms, TeaLeafError
    from exploringsycl_b200._lib import lib
    comms = Comms(session, rank, world, host_only=True)
    t0 = time.perf_counter()
    if rank == 1:
        time.sleep(0.3)
        lib().tl_comms_abort(comms.handle)  # abort_comms(), comms.c:79-82, then the process exits without a barrier
        out.put((rank, "aborted", 0.0))
        return
    try:
        comms.barrier()  # rank 1 never arrives
        out.put((rank, "no error", time.perf_counter() - t0))
    except TeaLeafError as e:
        out.put((rank, str(e), time.perf_counter() - t0))
    # the peer is gone: no collective teardown


def test_abort_poisons_the_segment_and_peers_fail_at_once():
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    session = "pytest_abort_%d" % os.getpid()
    procs = [ctx.Process(target=_abort_worker, args=(r, 2, session, out)) for r in range(2)]
    for p in procs:
        p.start()
    res = {r: (msg, dt) for r, msg, dt in (out.get(timeout=120) for _ in range(2))}
    for p in procs:
        p.join(timeout=60)
    assert "peer rank aborted" in res[0][0], res
    assert res[0][1] < 10.0  # not the 60 s time-out
    try:  # nobody ran the collective teardown: remove the segment
        os.unlink("/dev/shm/tl_b200_" + session)
    except OSError:
        pass
