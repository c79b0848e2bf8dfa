"""Parity of the CUDA path (through the C ABI, via the reference-named host functions) against
the CPU oracle on identical seeded inputs.  Tolerances: per-particle / per-cell arithmetic follows
the reference's evaluation order and is bit-exact where no summation order or FFT is involved;
density and everything downstream of the FFT is held to 1e-12 normalised L-infinity (the north
star's single-step bound)."""
import numpy as np
import pytest

from conftest import nlinf, make_particles

pytestmark = pytest.mark.gpu

TOL = 1e-12


@pytest.fixture(scope="module")
def PM(lib_built):
    from noname_b200 import ParticleMesh
    yield ParticleMesh
    ParticleMesh.clear_cache()


@pytest.fixture(scope="module")
def O():
    from oracle import pm_ref
    return pm_ref


DIMS = [((16, 16, 1), 2), ((12, 20, 1), 2), ((16, 16, 16), 3), ((8, 12, 20), 3), ((32, 32, 32), 3)]


@pytest.mark.parametrize("dims,rank", DIMS)
@pytest.mark.parametrize("interp", ["cic", "none"])
def test_deposit(PM, O, dims, rank, interp):
    x, _ = make_particles(1, 5000, rank, clustered=True)
    m = 0.37
    shape = (dims[2], dims[1], dims[0])
    ref = np.ones(shape)
    O.deposit(ref, x.copy(), m, interpolation=interp)
    got = np.ones(shape)
    PM.deposit(got, x, m, interpolation=interp)
    assert nlinf(got, ref) < TOL
    assert abs(got.sum() - m * x.shape[0]) < 1e-9 * m * x.shape[0]      # mass conservation


@pytest.mark.parametrize("dims,rank", [((16, 16, 1), 2), ((16, 16, 16), 3)])
def test_deposit_fixedpoint_bitexact_run_to_run(PM, O, dims, rank):
    x, _ = make_particles(2, 20000, rank, clustered=True)
    shape = (dims[2], dims[1], dims[0])
    PM.options["deposit_mode"] = "fixedpoint"
    try:
        outs = []
        for rep in range(3):
            perm = np.random.default_rng(rep).permutation(x.shape[0])     # order must not matter either
            got = np.empty(shape)
            PM.deposit(got, np.ascontiguousarray(x[perm]), 0.5, interpolation="cic")
            outs.append(got)
        assert np.array_equal(outs[0], outs[1]) and np.array_equal(outs[0], outs[2])
        ref = np.empty(shape)
        O.deposit(ref, x.copy(), 0.5, interpolation="cic")
        assert nlinf(outs[0], ref) < 1e-11      # 2^-44 quantisation per contribution
    finally:
        PM.options["deposit_mode"] = "atomic"


@pytest.mark.parametrize("dims,rank", DIMS)
def test_potential(PM, O, dims, rank):
    rng = np.random.default_rng(3)
    shape = (dims[2], dims[1], dims[0])
    rho = rng.random(shape)
    rho -= rho.mean()
    ref = np.empty(shape)
    O.compute_potential(ref, rho, np.zeros((shape[0], shape[1], shape[2] // 2 + 1), dtype=np.complex128))
    got = np.empty(shape)
    PM.compute_potential(got, rho, None)
    assert nlinf(got, ref) < TOL


@pytest.mark.parametrize("dims", [(16, 8, 64), (12, 10, 128), (32, 16, 256), (8, 8, 512), (8, 4, 1024), (66, 40, 256), (128, 128, 128)])
def test_potential_fused_z_pass(lib_built, O, dims):
    """the hand-written fused z-FFT * Green * z-iFFT kernel (N3 = 64 .. 1024) against the oracle and
    against the cuFFT-composed z path."""
    from noname_b200 import DevicePM
    rng = np.random.default_rng(33)
    shape = (dims[2], dims[1], dims[0])
    rho = rng.random(shape)
    ref = np.empty(shape)
    O.compute_potential(ref, rho - rho.mean(), np.zeros((shape[0], shape[1], shape[2] // 2 + 1), dtype=np.complex128))
    out = {}
    for zf in (3, 2, 0):          # 3 = split exchange (load overlaps the whole transform), 2 = single buffer, 0 = cuFFT z + Green kernel
        with DevicePM(dims, 1, rank=3) as dev:
            dev.set_option("zfused", zf)
            for rep in range(2):
                dev.set_field("rho", rho)
                dev.solve()
                out[zf] = dev.get_field("phi")
    assert nlinf(out[3], ref) < TOL
    assert nlinf(out[2], ref) < TOL
    assert nlinf(out[0], ref) < TOL
    assert nlinf(out[3], out[2]) < 1e-14    # same algorithm, different data movement (FMA contraction may differ)


@pytest.mark.parametrize("dims,rank", [((48, 20, 1), 2), ((64, 64, 1), 2), ((40, 24, 20), 3), ((64, 32, 32), 3), ((16, 16, 16), 3)])
@pytest.mark.parametrize("mode", ["atomic", "fixedpoint"])
def test_tiled_deposit_matches_oracle_and_global_path(lib_built, O, dims, rank, mode):
    """shared-memory tiled scatter on tile-sorted particles == oracle == global-atomic path, also after
    the particles strayed from their tiles (no re-sort); fixed-point sums are bit-identical."""
    from noname_b200 import DevicePM
    npart = 30000
    x, v = make_particles(21, npart, rank, clustered=True)
    m = 0.41
    shape = (dims[2], dims[1], dims[0])
    with DevicePM(dims, npart, rank=rank, deposit_mode=mode) as dev:
        dev.upload(x, v, m)
        dev.sort()
        for stray_dt in (0.0, 0.4, 3.0):           # 0.05-sigma velocities: up to ~cells of straying
            if stray_dt:
                dev.drift(stray_dt)
                dev.make_periodic()
            dev.set_option("tile_deposit", 1)
            dev.deposit()
            rt = dev.get_field("rho")
            dev.set_option("tile_deposit", 0)
            dev.deposit()
            rg = dev.get_field("rho")
            xr = x.copy()
            O.make_periodic(xr)
            for d in (0.4, 3.0):
                if stray_dt >= d:
                    O.drift(xr, v, d)
                    O.make_periodic(xr)
            ref = np.zeros(shape)
            O.deposit(ref, xr, m, interpolation="cic")
            assert nlinf(rt, ref) < TOL and nlinf(rg, ref) < TOL
            assert rt.sum() == pytest.approx(m * npart, rel=1e-12)
            if mode == "fixedpoint":
                assert np.array_equal(rt, rg)


@pytest.mark.parametrize("dims", [(24, 64, 8), (20, 128, 64), (10, 256, 12), (6, 512, 128), (8, 1024, 4), (64, 64, 64),
                                  (128, 128, 1), (130, 64, 1), (256, 64, 16), (512, 64, 4), (1024, 64, 2), (2048, 64, 1),
                                  (128, 20, 12), (1024, 12, 1)])
def test_potential_own_x_and_y_passes(lib_built, O, dims):
    """the hand-written 2-D transform passes against the oracle and against the cuFFT compositions:
    x = bulk-copy staged row FFT (r2c / c2r through a half-length complex transform, N1 = 128 .. 2048),
    y = TMA-staged column FFT (N2 = 64 .. 1024; icn = N1/2+1 is never a multiple of the 4-column group, so the
    out-of-bounds clipping of the last group is always exercised).  Row counts that are not a multiple of the
    4-row group exercise the ragged tail."""
    from noname_b200 import DevicePM
    rng = np.random.default_rng(35)
    shape = (dims[2], dims[1], dims[0])
    rho = rng.random(shape)
    ref = np.empty(shape)
    O.compute_potential(ref, rho - rho.mean(), np.zeros((shape[0], shape[1], shape[2] // 2 + 1), dtype=np.complex128))
    out = {}
    for key, (rf, cf) in {"own": (1, 1), "cufft_x": (0, 1), "cufft": (0, 0)}.items():
        with DevicePM(dims, 1, rank=3 if dims[2] > 1 else 2) as dev:
            dev.set_option("rowfft", rf)
            dev.set_option("colfft", cf)
            dev.set_field("rho", rho)
            dev.solve()
            out[key] = dev.get_field("phi")
            dev.set_field("rho", rho)          # a second solve through the same context (buffer reuse)
            dev.solve()
            assert np.array_equal(out[key], dev.get_field("phi"))
    for key in out:
        assert nlinf(out[key], ref) < TOL, key


@pytest.mark.parametrize("dims,block", [((128, 128, 8), 4), ((128, 128, 1), 4), ((256, 256, 6), 4), ((512, 512, 3), 4),
                                        ((1024, 1024, 4), 2), ((1024, 1024, 2), 1), ((128, 128, 64), 8)])
def test_potential_fused_2d_transform(lib_built, O, dims, block):
    """k_fft2d: x and y passes of a plane in one persistent kernel (work queue of row / column items, per-plane
    completion counters, intermediate in L2) == the separate-kernel path bit for bit (same arithmetic, different
    schedule) == oracle; 2-D and 3-D meshes, plane counts that are not a multiple of the block."""
    from noname_b200 import DevicePM, NNPMError
    with DevicePM((64, 64, 1), 1, rank=2) as probe:
        try:
            probe.set_option("fft2d", 1)
        except NNPMError as ex:   # the shipped library refuses the option: the kernel is built with `make EXPERIMENTS=1` only
            assert "NNPM_EXPERIMENTS" in str(ex)
            pytest.skip("k_fft2d is only built with -DNNPM_EXPERIMENTS (measured slower than the separate passes: DESIGN.md)")
    rng = np.random.default_rng(41)
    shape = (dims[2], dims[1], dims[0])
    rho = rng.random(shape)
    ref = np.empty(shape)
    O.compute_potential(ref, rho - rho.mean(), np.zeros((shape[0], shape[1], shape[2] // 2 + 1), dtype=np.complex128))
    out = {}
    for key, f2 in (("fused", 1), ("separate", 0)):
        with DevicePM(dims, 1, rank=3 if dims[2] > 1 else 2) as dev:
            dev.set_option("fft2d", f2)
            dev.set_option("fft2d_block", block)
            for rep in range(3):                       # repeated solves: queue / counters are reset every launch
                dev.set_field("rho", rho)
                dev.solve()
                out[key] = dev.get_field("phi")
    assert nlinf(out["fused"], ref) < TOL and nlinf(out["separate"], ref) < TOL
    assert np.array_equal(out["fused"], out["separate"])


@pytest.mark.parametrize("dims", [(12, 10, 128), (32, 16, 256), (8, 8, 512), (10, 4, 1024), (8, 4, 2048), (34, 6, 2048)])
def test_potential_fused_z_pass_radix2_wrapper(lib_built, O, dims):
    """N3 = 2 x (two-pass length): the persistent fused z pass with one radix-2 step around the two-pass engine
    (even / odd bins as two half-length transforms, parities recombined by a warp shuffle).  N3 = 2048 (the fine
    mesh of BASELINE config 5) always takes it; N3 <= 1024 through the "zsplit" hook.  Against the oracle and
    the cuFFT-composed z path; icn is never a multiple of the 4-column group (clipped last group)."""
    from noname_b200 import DevicePM
    rng = np.random.default_rng(37)
    shape = (dims[2], dims[1], dims[0])
    rho = rng.random(shape)
    ref = np.empty(shape)
    O.compute_potential(ref, rho - rho.mean(), np.zeros((shape[0], shape[1], shape[2] // 2 + 1), dtype=np.complex128))
    out = {}
    for zf in (2, 0):
        with DevicePM(dims, 1, rank=3) as dev:
            dev.set_option("zfused", zf)
            dev.set_option("zsplit", 1)
            for rep in range(2):
                dev.set_field("rho", rho)
                dev.solve()
                out[zf] = dev.get_field("phi")
    assert nlinf(out[0], ref) < TOL
    assert nlinf(out[2], ref) < TOL
    assert nlinf(out[2], out[0]) < 1e-13


def test_potential_ignores_mean(PM, O):
    """evolvePM subtracts mean(rho) (:768-769); the DC bin is zeroed (:501) so it must not matter."""
    rng = np.random.default_rng(4)
    rho = rng.random((16, 16, 16))
    a = np.empty_like(rho)
    b = np.empty_like(rho)
    PM.compute_potential(a, rho, None)
    PM.compute_potential(b, rho - rho.mean(), None)
    assert nlinf(a, b) < TOL


@pytest.mark.parametrize("dims,rank", DIMS)
def test_acceleration_bitexact(PM, O, dims, rank):
    rng = np.random.default_rng(5)
    shape = (dims[2], dims[1], dims[0])
    phi = rng.standard_normal(shape)
    ref = np.zeros(shape + (3,))
    O.compute_acceleration(ref, phi)
    got = np.full(shape + (3,), 7.0)
    PM.compute_acceleration(got, phi)
    assert np.array_equal(got, ref)


@pytest.mark.parametrize("dims,rank", DIMS)
@pytest.mark.parametrize("interp", ["cic", "none"])
def test_interp_bitexact(PM, O, dims, rank, interp):
    rng = np.random.default_rng(6)
    shape = (dims[2], dims[1], dims[0])
    acc = rng.standard_normal(shape + (3,))
    x, _ = make_particles(7, 3000, rank)
    ref = np.zeros((x.shape[0], rank))
    O.interpVecToPoints(ref, acc, x, interpolation=interp)
    got = np.zeros_like(ref)
    PM.interpVecToPoints(got, acc, x, interpolation=interp)
    assert np.array_equal(got, ref)


def test_interp_bugcompat_literal(PM, O):
    """ParticleMesh.jl:706 literally (k = floor(x3)*Ng3 + 1), SURVEY.md F7."""
    rng = np.random.default_rng(8)
    acc = rng.standard_normal((8, 8, 8, 3))
    x = rng.random((500, 3)) * 0.999
    ref = np.zeros((500, 3))
    O.interpVecToPoints(ref, acc, x, interpolation="cic", bugcompat=True)
    PM.options["bugcompat_interp_z"] = True
    try:
        got = np.zeros_like(ref)
        PM.interpVecToPoints(got, acc, x, interpolation="cic")
    finally:
        PM.options["bugcompat_interp_z"] = False
    assert np.array_equal(got, ref)


@pytest.mark.parametrize("rank", [2, 3])
def test_drift_kick_periodic_bitexact(PM, O, rank):
    x, v = make_particles(9, 4000, rank)
    pa = np.random.default_rng(10).standard_normal(x.shape)
    x[10] = -0.25
    x[11] = 1.25
    xr, vr = x.copy(), v.copy()
    O.drift(xr, vr, 0.013)
    O.make_periodic(xr)
    O.kick(vr, pa, 0.02)
    O.kick(vr, pa, 0.02, 1.3, 0.7)
    xg, vg = x.copy(), v.copy()
    PM.drift(xg, vg, 0.013)
    PM.make_periodic(xg)
    PM.kick(vg, pa, 0.02)
    PM.kick(vg, pa, 0.02, 1.3, 0.7)
    assert np.array_equal(xg, xr)
    assert np.array_equal(vg, vr)


def _one_step(dev_kw, O, dims, rank, npart, comoving, seed=11, sort=False):
    from noname_b200 import DevicePM
    x, v = make_particles(seed, npart, rank, clustered=True)
    m = float(np.prod(dims)) / npart
    f = O.Fields(dims, npart, rank)
    xr, vr = x.copy(), v.copy()
    dt = 0.01
    sc = (1.01, 1.02, 0.9, 1.03)
    if comoving:
        O.step_comoving(f, xr, vr, m, dt, *sc)
    else:
        O.step_static(f, xr, vr, m, dt)
    with DevicePM(dims, npart, rank=rank, **dev_kw) as dev:
        dev.upload(x, v, m)
        if sort:
            dev.sort()
        st = dev.step_comoving(dt, *sc) if comoving else dev.step_static(dt)
        phi = dev.get_field("phi")
        xg, vg = dev.download()
    return (xr, vr, f), (xg, vg, phi, st)


@pytest.mark.parametrize("dims,rank", [((16, 16, 1), 2), ((16, 16, 16), 3), ((8, 12, 20), 3)])
@pytest.mark.parametrize("comoving", [False, True])
@pytest.mark.parametrize("sort", [False, True])
def test_fused_step_matches_oracle(lib_built, O, dims, rank, comoving, sort):
    (xr, vr, f), (xg, vg, phi, st) = _one_step({}, O, dims, rank, 6000, comoving, sort=sort)
    assert nlinf(phi, f.phi) < TOL
    assert nlinf(vg, vr) < TOL
    assert np.max(np.abs(xg - xr)) < 1e-13
    assert abs(st.max_abs_v - np.max(np.abs(vr))) <= 1e-12 * np.max(np.abs(vr))
    assert abs(st.max_v - np.max(vr)) <= 1e-12 * abs(np.max(vr))
    assert abs(st.max_pa - np.max(f.pa)) <= 1e-10 * abs(np.max(f.pa))
    assert st.kernel_launches > 0 and st.n_halo_violations == 0


# kernel options of the tiled paths: bulk L2 prefetch of the work item's particle runs on / off, per-cell flush
PUSH_VARIANTS = [{}, {"l2pf_push": 0, "l2pf_deposit": 0}, {"bulk_flush": 0}]


# (128, 32, 32) and (128, 64, 1) have tiles whose phi box lies inside the mesh (staged by one TMA tensor
# copy); on the smaller meshes every box wraps periodically (cp.async rows)
@pytest.mark.parametrize("dims,rank", [((48, 20, 1), 2), ((64, 64, 1), 2), ((128, 64, 1), 2), ((40, 24, 20), 3),
                                       ((64, 32, 32), 3), ((128, 32, 32), 3)])
@pytest.mark.parametrize("comoving", [False, True])
@pytest.mark.parametrize("variant", PUSH_VARIANTS, ids=lambda o: "-".join(f"{k}{v}" for k, v in o.items()) or "default")
def test_tiled_push_multi_step_matches_oracle(lib_built, O, dims, rank, comoving, variant):
    """Tile-sorted shared-memory path over several steps (particles stray from their tiles between
    sorts and take the global fallback): same trajectory as the oracle."""
    from noname_b200 import DevicePM
    npart = 30000
    x, v = make_particles(77, npart, rank, clustered=True)
    v *= 4.0            # fast particles: many leave their tile box between sorts
    m = float(np.prod(dims)) / npart
    f = O.Fields(dims, npart, rank)
    xr, vr = x.copy(), v.copy()
    untiled = 0
    with DevicePM(dims, npart, rank=rank, sort_every=4) as dev:
        for k, val in variant.items():
            dev.set_option(k, val)
        dev.upload(x, v, m)
        for it in range(6):
            dt = 0.02
            sc = (1.0 + 0.01 * it, 1.01 + 0.01 * it, 0.8, 1.02 + 0.01 * it)
            if comoving:
                O.step_comoving(f, xr, vr, m, dt, *sc)
                st = dev.step_comoving(dt, *sc)
            else:
                O.step_static(f, xr, vr, m, dt)
                st = dev.step_static(dt)
            untiled += st.n_untiled
            assert abs(st.max_abs_v - np.max(np.abs(vr))) <= 1e-11 * np.max(np.abs(vr))
        phi = dev.get_field("phi")
        xg, vg = dev.download()
    assert untiled > 0                       # the fallback path was exercised
    assert nlinf(phi, f.phi) < 1e-11
    assert nlinf(vg, vr) < 1e-10
    assert np.max(np.abs(xg - xr)) < 1e-11


def test_fused_step_single_step_fields_rho_acc(lib_built, O):
    """north star: single-step density and acceleration fields within 1e-12 relative."""
    from noname_b200 import DevicePM
    dims, rank, npart = (32, 32, 32), 3, 32 ** 3
    x, v = O.synthetic_zeldovich((32, 32, 32), 3, 12345, 16, 4, 0.3 / 32, 0.1)
    m = 1.0
    f = O.Fields(dims, npart, rank)
    xr = x.copy()
    O.make_periodic(xr)
    O.deposit(f.rho, xr, m, "cic")
    O.compute_potential(f.phi, f.rho - f.rho.mean(), f.rt)
    O.compute_acceleration(f.acc, f.phi)
    with DevicePM(dims, npart, rank=rank) as dev:
        dev.upload(x, v, m)
        dev.make_periodic()
        dev.deposit()
        rho = dev.get_field("rho")
        dev.solve()
        dev.accel()
        acc = dev.get_field("acc")
    assert nlinf(rho, f.rho) < TOL
    assert nlinf(acc, f.acc) < TOL


def test_synthetic_ics_match_oracle(lib_built, O):
    from noname_b200 import DevicePM
    for pdims, dims, rank in (((16, 16, 16), (16, 16, 16), 3), ((24, 24, 1), (16, 16, 1), 2)):
        npart = int(np.prod(pdims))
        xr, vr = O.synthetic_zeldovich(pdims, rank, 777, 12, 3, 0.02, 0.5)
        with DevicePM(dims, npart, rank=rank) as dev:
            dev.init_synthetic(pdims, seed=777, nmodes=12, kmax=3, disp_rms=0.02, vfac=0.5, m=1.0)
            xg, vg = dev.download()
        assert np.max(np.abs(xg - xr)) < 1e-13
        assert np.max(np.abs(vg - vr)) < 1e-13


@pytest.mark.parametrize("mode", ["atomic", "fixedpoint"])
@pytest.mark.parametrize("variant", [{}, {"l2pf_push": 0, "l2pf_deposit": 0}], ids=["default", "nol2pf"])
def test_clustered_state_matches_oracle(lib_built, O, mode, variant):
    """BASELINE config 4 stand-in at test size: device-generated clustered particles equal the oracle's
    generator; a tile then holds many chunks of the work list (collapsed blobs), and the tiled deposit /
    push on that state equal the oracle."""
    from noname_b200 import DevicePM
    dims, n = (32, 32, 32), 48
    xr, vr = O.synthetic_clustered((n, n, n), 3, 777, 3, 0.5 / 32, 0.002)
    m = 32.0 ** 3 / n ** 3
    with DevicePM(dims, n ** 3, rank=3, sort_every=1, deposit_mode=mode) as dev:
        for k, val in variant.items():
            dev.set_option(k, val)
        dev.init_synthetic((n, n, n), kind="clustered", seed=777, nmodes=3, kmax=1, disp_rms=0.5 / 32, vfac=0.002, m=m)
        xg, vg = dev.download()
        assert np.max(np.abs(xg - xr)) < 1e-13 and np.max(np.abs(vg - vr)) < 1e-15
        dev.upload(xr, vr, m)                      # identical inputs from here on
        f = O.Fields(dims, n ** 3, 3)
        rho_ref = np.zeros((32, 32, 32))
        O.deposit(rho_ref, xr.copy(), m, interpolation="cic")
        assert rho_ref.max() > 1000 * rho_ref.mean()   # genuinely clustered: > 10^3 x the mean occupancy
        for it in range(3):
            sc = (1.0 + 0.01 * it, 1.02, 0.9, 1.03)
            O.step_comoving(f, xr, vr, m, 0.01, *sc)
            st = dev.step_comoving(0.01, *sc)
            assert st.n_halo_violations == 0
        assert nlinf(dev.get_field("phi"), f.phi) < 1e-11
        xg, vg = dev.download()
        assert np.max(np.abs(xg - xr)) < 1e-11 and nlinf(vg, vr) < 1e-10
        dev.deposit()
        ref = np.zeros((32, 32, 32))
        O.deposit(ref, xr.copy(), m, interpolation="cic")
        assert nlinf(dev.get_field("rho"), ref) < TOL


@pytest.mark.parametrize("sort_every", [0, 1])
def test_edge_positions_pileups_and_empty_input(lib_built, O, sort_every):
    """Domain edge cases through the fused step, tiled (sorted) and global (unsorted) kernels alike:
    coordinates exactly 0.0, 1.0 (make_periodic leaves it, ParticleMesh.jl:411: cell index N folds to 0),
    the largest double below 1, a denormal; 300 particles on one point (every atomic of a warp on the same
    cells); a context with zero particles."""
    import itertools
    from noname_b200 import DevicePM
    dims = (64, 32, 32)
    edge = [0.0, 1.0, np.nextafter(1.0, 0.0), 5e-324, 0.5]
    pts = np.array(list(itertools.product(edge, repeat=3)))
    rng = np.random.default_rng(3)
    x = np.concatenate([rng.random((2000, 3)), np.tile([[0.5, 0.5, 0.5]], (300, 1)),
                        np.tile([[0.999, 0.999, 0.999]], (300, 1)), pts])
    v = (rng.random(x.shape) - 0.5) * 0.02
    m = 0.7
    npart = len(x)
    shape = (dims[2], dims[1], dims[0])
    f = O.Fields(dims, npart, 3)
    xr, vr = x.copy(), v.copy()
    sc = (1.0, 1.01, 0.8, 1.02)
    with DevicePM(dims, npart, rank=3, sort_every=sort_every) as dev:
        dev.upload(x, v, m)
        if sort_every:
            dev.sort()
        dev.deposit()
        ref = np.zeros(shape)
        xw = x.copy()
        O.make_periodic(xw)
        O.deposit(ref, xw, m, interpolation="cic")
        rho = dev.get_field("rho")
        assert nlinf(rho, ref) < TOL
        assert rho.sum() == pytest.approx(m * npart, rel=1e-12)
        for it in range(2):
            O.step_comoving(f, xr, vr, m, 0.01, *sc)
            st = dev.step_comoving(0.01, *sc)
            assert st.n_halo_violations == 0
        xg, vg = dev.download()
    assert np.max(np.abs(xg - xr)) < 1e-12 and nlinf(vg, vr) < 1e-11
    assert xg.min() >= 0.0 and xg.max() <= 1.0
    # zero particles: the step runs, the fields are zero, the reductions report an empty set
    with DevicePM(dims, 16, rank=3, sort_every=sort_every) as dev:
        dev.upload(np.zeros((0, 3)), np.zeros((0, 3)), 1.0)
        st = dev.step_comoving(0.01, *sc)
        assert st.npart_local == 0 and st.max_abs_v == 0.0
        assert not dev.get_field("phi").any()
        xg, vg = dev.download()
        assert xg.shape == (0, 3) and vg.shape == (0, 3)


def test_sort_keeps_particles_and_restores_order(lib_built):
    from noname_b200 import DevicePM
    x, v = make_particles(21, 10000, 3)
    with DevicePM((16, 16, 16), 10000, rank=3) as dev:
        dev.upload(x, v, 1.0)
        dev.sort()
        xs, vs, ids = dev.download(ids=True)
        assert sorted(ids.tolist()) == list(range(10000))
        assert np.array_equal(xs, x[ids]) and np.array_equal(vs, v[ids])
        xo, vo = dev.download()
        assert np.array_equal(xo, x) and np.array_equal(vo, v)


@pytest.mark.parametrize("cells", [1, 0])
def test_sort_orders_by_cell_and_survives_a_nonzero_first_id(lib_built, cells):
    """cell-order sort: inside every tile the particles come out ordered by (k, j, i) cell, tiles in tile order; a single-GPU
    upload whose ids start at first_id != 0 is still restored to its original order (out-of-bounds permute)"""
    from noname_b200 import DevicePM
    dims = (64, 32, 32)
    x, v = make_particles(23, 30000, 3, clustered=True)
    with DevicePM(dims, 30000, rank=3) as dev:
        dev.set_option("sort_cells", cells)
        dev.upload(x, v, 1.0, first_id=5000)
        dev.sort()
        xs, vs, ids = dev.download(ids=True)
        assert sorted(ids.tolist()) == list(range(5000, 35000))
        assert np.array_equal(xs, x[ids - 5000]) and np.array_equal(vs, v[ids - 5000])
        xw = np.where(xs > 1.0, xs - 1.0, np.where(xs < 0.0, xs + 1.0, xs))
        cell = [np.minimum(np.floor(xw[:, d] * dims[d]).astype(np.int64) % dims[d], dims[d] - 1) for d in range(3)]
        tile = (cell[0] // 32) + 2 * ((cell[1] // 8) + 4 * (cell[2] // 8))
        assert np.all(np.diff(tile) >= 0)                               # tile-sorted either way
        if cells:
            key = tile * 2048 + (cell[0] % 32) + 32 * ((cell[1] % 8) + 8 * (cell[2] % 8))
            assert np.all(np.diff(key) >= 0)                            # and cell-ordered inside the tiles
        xo, vo = dev.download()
        assert np.array_equal(xo, x) and np.array_equal(vo, v)


@pytest.mark.parametrize("dims,rank", [((16, 16, 1), 2), ((16, 16, 16), 3), ((8, 12, 20), 3)])
def test_power_spectrum(lib_built, O, dims, rank):
    from noname_b200 import DevicePM
    x, v = make_particles(31, 20000, rank, clustered=True)
    kr, pr = O.powerSpectrum(x.copy(), 1.0, dims, "cic")
    with DevicePM(dims, x.shape[0], rank=rank) as dev:
        dev.upload(x, v, 1.0)
        kg, pg = dev.power_spectrum()
    assert np.allclose(kg, kr, rtol=1e-15)
    ok = np.isfinite(pr)
    assert np.array_equal(ok, np.isfinite(pg))
    assert np.max(np.abs(pg[ok] / pr[ok] - 1.0)) < 1e-10


def test_nstep_evolution_pk_within_1e6(lib_built, O):
    """north star: final P(k) after N steps within 1e-6 (comoving loop, dt control included)."""
    from noname_b200 import ParticleMesh as PM, cosmology as cos
    dims, pdims, rank = (16, 16, 16), (16, 16, 16), 3
    cosmo_kw = dict(h=0.7, OmegaM=0.3, OmegaR=0.0)
    oc = O.Cosmo(**cosmo_kw)
    zi, zf = 50.0, 20.0
    ts, te = O.start_stop_times(oc, zi, zf)
    a0 = O.a_from_t_internal(oc, ts, zi, zwherea1=zi)
    ad0 = O.dadt_from_t_internal(oc, ts, zi)
    x0, v0 = O.synthetic_zeldovich(pdims, rank, 5, 16, 3, 0.4 / 16, ad0 / a0)
    m = 0.3
    conf = dict(StartTime=ts, StopTime=te, StartCycle=0, StopCycle=12, InitialDt=1e-4, CourantFactor=0.5,
                MaximumDt=1.0, DtFractionOfTime=0.1, InitialRedshift=zi)
    xr, vr = x0.copy(), v0.copy()
    logr = []
    O.evolvePMCosmology(xr, vr, m, dims, dict(conf), oc, log=logr)
    kr, pr = O.powerSpectrum(xr.copy(), m, dims, "cic")

    from noname_b200.GridTypes import GridPatch, GridData
    gp = {1: GridPatch([0, 0, 0], [1, 1, 1], 1, 1, dims)}
    gd = {1: GridData({"ρD": np.ones(dims[::-1]), "Φ": np.ones(dims[::-1])})}
    p = {"x": x0.copy(), "v": v0.copy(), "m": m}
    c = dict(conf, ParticleDimensions=list(pdims), ParticleDepositInterpolation="cic", ParticleBackInterpolation="cic",
             OutputEveryNCycle=0, OutputDtDataDump=0.0, LastOutputTime=0.0, OutputDisabled=True, _rank=3,
             _cosmo=cos.Cosmology(**cosmo_kw))
    logg = []
    PM.evolvePMCosmology(gp, gd, p, c, log=logg)
    assert len(logg) == len(logr) == 12
    assert abs(logg[-1][0] - logr[-1][0]) < 1e-12 * logr[-1][0]
    kg, pg = PM.powerSpectrum(gp, gd, p, c)
    ok = np.isfinite(pr)
    assert np.max(np.abs(pg[ok] / pr[ok] - 1.0)) < 1e-6
    assert np.max(np.abs(p["x"] - xr)) < 1e-9


def test_error_behaviour(PM):
    import noname_b200 as nb
    with pytest.raises(ValueError):
        PM.deposit(np.zeros((4, 4, 4)), np.zeros((3, 3)), 1.0, interpolation="tsc")
    with pytest.raises(TypeError):
        PM.deposit(np.zeros((4, 4, 4), dtype=np.float32), np.zeros((3, 3)), 1.0, interpolation="cic")
    with nb.DevicePM((8, 8, 8), 10) as dev:
        with pytest.raises(nb.NNPMError):
            dev.solve()                     # nothing deposited yet
        with pytest.raises(nb.NNPMError):
            dev.upload(np.zeros((11, 3)), None, 1.0)   # over capacity
