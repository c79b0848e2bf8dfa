import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def nlinf(a, b):
    """normalised L-infinity difference max|a-b| / max|b| (the parity norm of DESIGN.md)."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    den = np.max(np.abs(b))
    return float(np.max(np.abs(a - b)) / (den if den > 0 else 1.0))


@pytest.fixture(scope="session")
def lib_built():
    from noname_b200 import _lib
    if not os.path.exists(_lib.LIB_PATH):
        import __graft_entry__
        __graft_entry__.build()
    return _lib.load()


def make_particles(seed, npart, rank, clustered=False):
    """seeded test particles in [0,1) plus edge cases (0.0, 1.0 - eps, exact cell boundaries)."""
    rng = np.random.default_rng(seed)
    x = rng.random((npart, rank))
    if clustered:
        nb = max(1, npart // 3)
        c = rng.random((4, rank))
        x[:nb] = np.mod(c[rng.integers(0, 4, nb)] + 0.01 * rng.standard_normal((nb, rank)), 1.0)
    v = 0.05 * rng.standard_normal((npart, rank))
    if npart >= 8:
        x[0] = 0.0
        x[1] = np.nextafter(1.0, 0.0)
        x[2] = 0.5
        x[3] = 1.0 / 8.0
        x[4] = 1.0            # make_periodic leaves x == 1.0 alone (ParticleMesh.jl:411-412)
    return np.ascontiguousarray(x), np.ascontiguousarray(v)
