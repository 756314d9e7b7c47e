"""Shared helpers for the parity tests (oracle = checker, CUDA library = thing under test)."""
import ctypes as C
import os

import numpy as np

from oracle import oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DECKS = os.path.join(ROOT, "tests", "decks")
GOLDEN = os.path.join(ROOT, "tests", "golden")
HD = 2


def rng_fields(nx, ny, seed=1234, hd=HD):
    """Deterministic random dense fields (p,u,r in [-1,1), kx,ky in [0.1,1))."""
    r = np.random.default_rng(seed)
    x, y = nx + 2 * hd, ny + 2 * hd
    f = {}
    for name in ("p", "u", "r", "w", "u0", "sd"):
        f[name] = r.uniform(-1.0, 1.0, (y, x))
    for name in ("kx", "ky"):
        f[name] = r.uniform(0.1, 1.0, (y, x))
    f["density"] = r.uniform(0.1, 100.0, (y, x))
    f["energy"] = r.uniform(1e-4, 25.0, (y, x))
    f["energy0"] = r.uniform(1e-4, 25.0, (y, x))
    f["volume"] = np.full((y, x), 0.01)
    return f


FIELD_IDS = dict(density=0, energy0=1, energy=2, u=3, p=4, sd=5, u0=6, r=7, w=8, kx=9, ky=10, volume=11)


def upload(chunk, fields):
    for name, arr in fields.items():
        chunk.write(FIELD_IDS[name], arr)


def download(chunk, names):
    return {n: chunk.read(FIELD_IDS[n]) for n in names}


def dbl():
    return C.c_double(0.0)


def rel(a, b):
    return abs(a - b) / max(abs(a), abs(b), 1e-300)


def has_gpu():
    try:
        from exploringsycl_b200 import lib
        return lib().tl_device_count() > 0
    except Exception:
        return False
