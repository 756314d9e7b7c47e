"""Halo parity (bit-exact): reflective local halos, pack / unpack, message layout."""
import numpy as np
import pytest

from oracle import oracle as O
from tl_testutil import HD, rng_fields, upload

pytestmark = pytest.mark.gpu

SIZES = [(5, 3), (37, 23), (300, 129)]
EXCH = ["density", "energy0", "energy", "u", "p", "sd"]  # exchange index order, shared.h:40-45


@pytest.mark.parametrize("size", SIZES, ids=lambda s: "%dx%d" % s)
@pytest.mark.parametrize("depth", [1, 2])
@pytest.mark.parametrize("nbs", [(-1, -1, -1, -1), (3, -1, -1, 5), (-1, 2, 7, -1), (1, 2, 3, 4)])
def test_local_halos(size, depth, nbs):
    from exploringsycl_b200 import Chunk, Settings
    nx, ny = size
    x, y = nx + 2 * HD, ny + 2 * HD
    ch = Chunk(nx, ny, HD, 10, neighbours=nbs)
    f = rng_fields(nx, ny, seed=7)
    upload(ch, f)
    s = Settings()
    s.fields_to_exchange = [True, False, True, True, True, False]
    ch.run_local_halos(s, depth)
    for idx, name in enumerate(EXCH):
        exp = f[name].copy()
        if s.fields_to_exchange[idx]:
            for face in (O.LEFT, O.RIGHT, O.TOP, O.BOTTOM):  # kernel_interface.cpp:112-118 order
                if nbs[face] == O.EXTERNAL:
                    O.lib().orc_local_halo(x, y, HD, depth, face, exp)
        assert np.array_equal(ch.read(idx), exp), name
    ch.close()


@pytest.mark.parametrize("size", SIZES, ids=lambda s: "%dx%d" % s)
@pytest.mark.parametrize("depth", [1, 2])
@pytest.mark.parametrize("face", [0, 1, 2, 3])
def test_pack_unpack_host_buffer(size, depth, face):
    """run_pack_or_unpack with a HOST buffer, one field at a time (kernel_interface.cpp:129-167)."""
    from exploringsycl_b200 import Chunk
    nx, ny = size
    x, y = nx + 2 * HD, ny + 2 * HD
    ch = Chunk(nx, ny, HD, 10)
    f = rng_fields(nx, ny, seed=11)
    upload(ch, f)
    n = depth * (y if face in (0, 1) else x)
    for fid, name in ((3, "u"), (0, "density")):
        buf = np.zeros(n)
        ch.run_pack_or_unpack(depth, face, True, fid, buf)
        exp = np.zeros(n)
        O.lib().orc_pack(x, y, HD, depth, face, f[name], exp)
        assert np.array_equal(buf, exp)
        msg = np.random.default_rng(face).uniform(-5, 5, n)
        ch.run_pack_or_unpack(depth, face, False, fid, msg)
        expf = f[name].copy()
        O.lib().orc_unpack(x, y, HD, depth, face, expf, msg)
        assert np.array_equal(ch.read(fid), expf)
    ch.close()


@pytest.mark.parametrize("depth", [1, 2])
@pytest.mark.parametrize("face", [0, 1, 2, 3])
def test_multi_field_message_layout(depth, face):
    """All flagged fields in one launch: concatenated in index order (remote_halo_driver.c:132-184)."""
    import ctypes as C
    from exploringsycl_b200 import Chunk
    from exploringsycl_b200._lib import check
    nx, ny = 61, 45
    x, y = nx + 2 * HD, ny + 2 * HD
    ch = Chunk(nx, ny, HD, 10)
    f = rng_fields(nx, ny, seed=13)
    upload(ch, f)
    flags = [1, 0, 1, 1, 1, 0]
    ln = C.c_int()
    check(ch.L.tl_pack_face_device(ch.handle, (C.c_int * 6)(*flags), depth, face, 1, C.byref(ln)))
    per = depth * (y if face in (0, 1) else x)
    assert ln.value == per * sum(flags)
    got = np.zeros(ln.value)
    check(ch.L.tl_face_buffer_read(ch.handle, face, 1, got, ln.value))
    exp = []
    for idx, name in enumerate(EXCH):
        if flags[idx]:
            b = np.zeros(per)
            O.lib().orc_pack(x, y, HD, depth, face, f[name], b)
            exp.append(b)
    assert np.array_equal(got, np.concatenate(exp))
    ch.close()
