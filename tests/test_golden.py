"""Committed golden fixtures (tests/golden/*.npz, written by tests/golden/make_golden.py).
CPU: the oracle and its C twin reproduce them (regression pin of the checker) and the closed-form
vectors.  GPU (-m gpu): the CUDA path, through the C ABI, reproduces them -- without importing
the oracle at all."""
import glob
import os

import numpy as np
import pytest

from conftest import nlinf

HERE = os.path.dirname(os.path.abspath(__file__))
STEP_CASES = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(HERE, "golden", "step*.npz")))
TOL = 1e-12


def load(name):
    return dict(np.load(os.path.join(HERE, "golden", name + ".npz")))


def test_fixtures_present():
    assert len(STEP_CASES) == 4 and os.path.exists(os.path.join(HERE, "golden", "analytic_16.npz"))


# ------------------------------------------------------------------------------- CPU: the checker
@pytest.mark.parametrize("name", STEP_CASES)
@pytest.mark.parametrize("impl", ["numpy", "c"])
def test_oracle_reproduces_golden(name, impl):
    from oracle import pm_ref as O
    from oracle.pm_ref_c import PMRefC
    g = load(name)
    dims, rank = tuple(int(d) for d in g["dims"]), int(g["rank"])
    eng = O if impl == "numpy" else PMRefC(threads=1)
    x, v, m = g["x0"].copy(), g["v0"].copy(), float(g["m"])
    f = O.Fields(dims, x.shape[0], rank)
    for s, dt in enumerate(g["dts"]):
        if int(g["comoving"]):
            eng.step_comoving(f, x, v, m, float(dt), *[float(q) for q in g["scalars"][s]])
        else:
            eng.step_static(f, x, v, m, float(dt))
        if s == 0:
            assert np.array_equal(f.rho, g["rho1"])
            assert nlinf(f.phi, g["phi1"]) < 1e-14 and nlinf(f.acc, g["acc1"]) < 1e-13
            assert nlinf(f.pa, g["pa1"]) < 1e-13
            assert np.max(np.abs(x - g["x1"])) < 1e-15 and nlinf(v, g["v1"]) < 1e-14
    assert np.max(np.abs(x - g["xN"])) < 1e-14 and nlinf(v, g["vN"]) < 1e-13


def test_oracle_reproduces_golden_initial_conditions():
    """seeded ICs (scale-free and multiplePeaksPS random fields, the pancake): same numbers as when the fixture was written
    -- pins the hash-based draws, the Hermitian planes and the spectrum windows of the checker of nnpm_init_grf[_peaks]"""
    import math
    import sys
    from oracle import pm_ref as O
    sys.path.insert(0, os.path.join(HERE, "golden"))
    from make_golden import IC_CASES
    g = load("ics")
    for name, (pd, rank, seed, n, rms, vfac, peaks) in IC_CASES.items():
        x, v = O.synthetic_grf(pd, rank, seed, n, rms, vfac, peaks=peaks)
        assert np.max(np.abs(x - g[name + "_x"])) < 1e-15 and np.max(np.abs(v - g[name + "_v"])) < 1e-15, name
        assert np.sqrt(np.mean((v / vfac) ** 2)) == pytest.approx(rms, rel=1e-12) if vfac else not v.any()
    x, v = O.zeldovich_pancake((16, 8, 1), 2, 10.0, 400.0, 1.0, math.sqrt(2.0 / 3.0))
    assert np.array_equal(x, g["pancake_x"]) and np.array_equal(v, g["pancake_v"])


def test_oracle_matches_analytic_vectors():
    from oracle import pm_ref as O
    g = load("analytic_16")
    phi = np.empty_like(g["rho"])
    O.compute_potential(phi, g["rho"], np.zeros((16, 16, 9), dtype=np.complex128))
    assert nlinf(phi, g["phi"]) < 1e-13
    acc = np.zeros((16, 16, 16, 3))
    O.compute_acceleration(acc, phi)
    assert nlinf(acc, g["acc"]) < 1e-12
    rho = np.zeros((16, 16, 16))
    O.deposit(rho, g["xp"].copy(), float(g["mp"]), interpolation="cic")
    assert np.max(np.abs(rho - g["wp"])) < 1e-15


# ------------------------------------------------------------------------------- GPU: the product
@pytest.mark.gpu
@pytest.mark.parametrize("name", STEP_CASES)
@pytest.mark.parametrize("sort_every", [0, 1])
def test_device_reproduces_golden_steps(lib_built, name, sort_every):
    from noname_b200 import DevicePM
    g = load(name)
    dims, rank = tuple(int(d) for d in g["dims"]), int(g["rank"])
    x0, v0, m = np.ascontiguousarray(g["x0"]), np.ascontiguousarray(g["v0"]), float(g["m"])
    with DevicePM(dims, x0.shape[0], rank=rank, sort_every=sort_every) as dev:
        dev.upload(x0, v0, m)
        for s, dt in enumerate(g["dts"]):
            if int(g["comoving"]):
                st = dev.step_comoving(float(dt), *[float(q) for q in g["scalars"][s]])
            else:
                st = dev.step_static(float(dt))
            assert st.kernel_launches > 0
            if s == 0:
                assert nlinf(dev.get_field("phi"), g["phi1"]) < TOL
                x1, v1 = dev.download()
                assert np.max(np.abs(x1 - g["x1"])) < 1e-13 and nlinf(v1, g["v1"]) < TOL
                assert st.max_abs_v == pytest.approx(np.max(np.abs(g["v1"])), rel=1e-11)
        xN, vN = dev.download()
        assert nlinf(dev.get_field("phi"), g["phiN"]) < 1e-11
        assert np.max(np.abs(xN - g["xN"])) < 1e-12 and nlinf(vN, g["vN"]) < 1e-11
        k, p = dev.power_spectrum()
        ok = np.isfinite(g["pk_p"])
        assert np.allclose(k, g["pk_k"], rtol=1e-14) and np.max(np.abs(p[ok] / g["pk_p"][ok] - 1)) < 1e-6


@pytest.mark.gpu
@pytest.mark.parametrize("name", STEP_CASES)
def test_device_reproduces_golden_fields(lib_built, name):
    """single-step density / acceleration fields (north star: 1e-12 relative in float64)"""
    from noname_b200 import DevicePM
    g = load(name)
    dims, rank = tuple(int(d) for d in g["dims"]), int(g["rank"])
    x0, v0, m = np.ascontiguousarray(g["x0"]), np.ascontiguousarray(g["v0"]), float(g["m"])
    with DevicePM(dims, x0.shape[0], rank=rank) as dev:
        dev.upload(x0, v0, m)
        if int(g["comoving"]):
            dev.drift(float(g["dts"][0]) / 2 / float(g["scalars"][0][0]))
        dev.make_periodic()
        dev.deposit()
        assert nlinf(dev.get_field("rho"), g["rho1"]) < TOL
        dev.solve()
        assert nlinf(dev.get_field("phi"), g["phi1"]) < TOL
        dev.accel()
        assert nlinf(dev.get_field("acc"), g["acc1"]) < TOL
        dev.interp()
        assert nlinf(dev.get_field("pa"), g["pa1"]) < TOL


@pytest.mark.gpu
def test_device_matches_analytic_vectors(lib_built):
    from noname_b200 import DevicePM
    g = load("analytic_16")
    with DevicePM((16, 16, 16), 1, rank=3) as dev:
        dev.set_field("rho", np.ascontiguousarray(g["rho"]))
        dev.solve()
        assert nlinf(dev.get_field("phi"), g["phi"]) < TOL
        dev.accel()
        assert nlinf(dev.get_field("acc"), g["acc"]) < TOL
        dev.upload(np.ascontiguousarray(g["xp"]), None, float(g["mp"]))
        dev.deposit()
        assert np.max(np.abs(dev.get_field("rho") - g["wp"])) < 1e-15
