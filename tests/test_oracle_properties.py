"""Property tests of the CPU oracle (hypothesis): the numpy restatement, its literal-loop twin and the C restatement
agree on ragged shapes, empty particle sets and positions on the periodic edges; and the restated operators keep the
invariants the formulas of src/ParticleMesh.jl imply (mass, linearity, constant-field interpolation, idempotent wrap)."""
import numpy as np
import pytest
from hypothesis import HealthCheck, given, settings, strategies as st

from oracle import pm_ref as O

SET = settings(max_examples=40, deadline=None, derandomize=True, database=None, suppress_health_check=[HealthCheck.too_slow, HealthCheck.function_scoped_fixture])


@st.composite
def problems(draw, max_part=40):
    rank = draw(st.sampled_from([2, 3]))
    n1, n2 = draw(st.integers(2, 9)), draw(st.integers(2, 9))
    n3 = draw(st.integers(2, 7)) if rank == 3 else 1
    npart = draw(st.integers(0, max_part))
    seed = draw(st.integers(0, 2 ** 31 - 1))
    rng = np.random.default_rng(seed)
    x = rng.random((npart, rank))
    # sprinkle the special positions of the periodic box: 0, 1 (make_periodic leaves it, :411), the last double below 1
    for row in range(min(npart, 3)):
        x[row, rng.integers(rank)] = [0.0, 1.0, np.nextafter(1.0, 0.0)][row]
    v = 0.1 * rng.standard_normal((npart, rank))
    return (n1, n2, n3), rank, x, v, float(draw(st.floats(0.1, 3.0)))


@pytest.fixture(scope="module")
def CR():
    from oracle.pm_ref_c import PMRefC
    return PMRefC(threads=1)


@SET
@given(problems())
def test_deposit_conserves_mass_and_matches_loops_and_c(CR, pr):
    dims, rank, x, v, m = pr
    shape = (dims[2], dims[1], dims[0])
    for interp in ("cic", "none"):
        a, b = np.full(shape, 7.0), np.full(shape, -3.0)            # stale contents must be overwritten (:290-292)
        O.deposit(a, x, m, interpolation=interp)
        CR.deposit(b, x, m, interp)
        assert np.array_equal(a, b)
        assert a.sum() == pytest.approx(m * x.shape[0], rel=1e-13, abs=1e-300)
        assert (a >= 0).all()
    c = np.zeros(shape)
    O.cicdensity_loops(c, x, m)
    O.deposit(a, x, m, interpolation="cic")
    assert np.array_equal(a, c)


@SET
@given(problems(), st.floats(0.25, 4.0))
def test_deposit_is_linear_in_the_mass(CR, pr, s):
    dims, rank, x, v, m = pr
    shape = (dims[2], dims[1], dims[0])
    a, b = np.zeros(shape), np.zeros(shape)
    O.deposit(a, x, m, interpolation="cic")
    O.deposit(b, x, m * s, interpolation="cic")
    assert np.allclose(b, a * s, rtol=1e-14, atol=0)


@SET
@given(problems())
def test_interpolating_a_constant_field_returns_the_constant(pr):
    """the 4 / 8 CIC weights sum to one (:676-721): a uniform acc comes back unchanged at every particle"""
    dims, rank, x, v, m = pr
    acc = np.empty((dims[2], dims[1], dims[0], 3))
    acc[...] = (0.5, -1.25, 2.0)
    pa = np.zeros((x.shape[0], rank))
    O.interpVecToPoints(pa, acc, x, interpolation="cic")
    assert np.allclose(pa, np.array([0.5, -1.25, 2.0])[:rank], rtol=1e-14, atol=0)


@SET
@given(problems())
def test_potential_ignores_the_mean_and_has_zero_mean(pr):
    """:501 zeroes the DC bin: phi(rho + c) == phi(rho), mean(phi) == 0"""
    dims, rank, x, v, m = pr
    if dims[0] != dims[1] or (rank == 3 and dims[2] != dims[0]):
        dims = (dims[0],) * 2 + ((dims[0],) if rank == 3 else (1,))      # the reference's 1/Nx^2 is for cubic meshes (:502)
    shape = (dims[2], dims[1], dims[0])
    rng = np.random.default_rng(x.shape[0])
    rho = rng.random(shape)
    f = O.Fields(dims, 1, rank)
    p1, p2 = np.empty(shape), np.empty(shape)
    O.compute_potential(p1, rho, f.rt)
    O.compute_potential(p2, rho + 3.5, f.rt)
    scale = max(np.abs(p1).max(), 1e-300)
    assert np.abs(p1 - p2).max() <= 1e-12 * scale and abs(p1.mean()) <= 1e-13 * scale


@SET
@given(st.lists(st.floats(-1.0, 2.0, allow_nan=False), min_size=0, max_size=30))
def test_make_periodic_is_a_single_wrap_and_idempotent_inside_the_box(vals):
    x = np.array(vals, dtype=np.float64)
    y = x.copy()
    O.make_periodic(y)
    assert ((y >= 0.0) & (y <= 1.0)).all()
    z = y.copy()
    O.make_periodic(z)
    assert np.array_equal(y, z)
    assert np.array_equal(y[(x >= 0) & (x <= 1)], x[(x >= 0) & (x <= 1)])


@SET
@given(problems(max_part=25), st.booleans())
def test_full_step_numpy_equals_c_on_ragged_problems(CR, pr, comoving):
    dims, rank, x, v, m = pr
    npart = x.shape[0]
    fa, fb = O.Fields(dims, npart, rank), O.Fields(dims, npart, rank)
    xa, va, xb, vb = x.copy(), v.copy(), x.copy(), v.copy()
    if comoving:
        sc = (1.01, 1.02, 0.8, 1.03)
        O.step_comoving(fa, xa, va, m, 0.01, *sc)
        CR.step_comoving(fb, xb, vb, m, 0.01, *sc)
    else:
        O.step_static(fa, xa, va, m, 0.01)
        CR.step_static(fb, xb, vb, m, 0.01)
    assert np.array_equal(fa.rho, fb.rho)
    sphi = max(np.abs(fa.phi).max(), 1e-300)
    assert np.abs(fa.phi - fb.phi).max() <= 1e-13 * sphi
    if npart:
        assert np.abs(xa - xb).max() <= 1e-13 and np.abs(va - vb).max() <= 1e-13 * max(1.0, np.abs(va).max())
