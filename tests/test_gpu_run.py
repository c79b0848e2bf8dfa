"""End-to-end parity through the product surface -- run_noname(conf) -> gp, gd, p -- on the reference's shipped
configurations, device-generated initial conditions against their oracle twins, and the north star's single-step
field tolerance at a BASELINE size (512^3, config 3) against the threaded C restatement."""
import os

import numpy as np
import pytest

from conftest import nlinf

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

# run/CosmologyPM/particle_mesh.conf of the reference, verbatim (the reference tree is not on the GPU box)
SHIPPED_COSMOLOGY_PM = """
verbose = false
Initializer = InitializeParticleSimulation # function for Problem initialization
#ProblemDefinition = UniformParticles # string identifier to use in Initializer routine
#ProblemDefinition = PowerSpectrumParticles # string identifier to use in Initializer routine
#InputPowerSpectrum = multiplePeaksPS(-1, 1, [0.1, 1, 10], 1.)
#InputPowerSpectrum = scaleFreePS(-1, .3)
#ProblemDefinition = UniformParticleWind # string identifier to use in Initializer routine
#UniformParticleWindVelocity = 1.
ProblemDefinition = ZeldovichPancakeParticles # string identifier to use in Initializer routine
ZeldovichCausticRedshift = 10.


Evolve = evolvePMCosmology
TopgridDimensions  = [64, 64, 1] # main grid dimensions (active zones)
ParticleDimensions = [128, 128, 1]

MaximumDt = 1.0
InitialDt = 0.0001

OutputPowerSpectrum = false

CourantFactor = 0.5
DtFractionOfTime = 0.1  # dt should not be larger than this fraction of Current Time

ParticleDepositInterpolation = cic # none and cic implemented so far
ParticleBackInterpolation    = cic # none and cic implemented so far

OutputDirPrefix   = "pm_"
OutputDtDataDump = 3.5
Cosmology = cosmology(h=1.,OmegaM=1.0,OmegaR=0.0) # see https://github.com/JuliaAstro/Cosmology.jl
OmegaBaryon = 0.
ComovingCoordinates = true
InitialRedshift = 400
FinalRedshift   = 9
"""

# run/PM/particle_mesh.conf of the reference, verbatim
SHIPPED_PM = """
verbose = false
Initializer = InitializeParticleSimulation # function for Problem initialization
#ProblemDefinition = UniformParticles # string identifier to use in Initializer routine
ProblemDefinition = PowerSpectrumParticles # string identifier to use in Initializer routine
InputPowerSpectrum = multiplePeaksPS(-1, 1, [0.1, 1, 10], 1.)
Evolve = evolvePM
TopgridDimensions  = [1000, 1000, 1] # main grid dimensions (active zones)
ParticleDimensions = [100, 100, 1]

MaximumDt = 0.1
StopTime = 1.
CourantFactor = 0.2

ParticleDepositInterpolation = cic # none and cic implemented so far
ParticleBackInterpolation    = cic # none and cic implemented so far

OutputDirPrefix   = "pm_"
OutputDtDataDump = 0.5
"""


@pytest.fixture(scope="module")
def O():
    from oracle import pm_ref
    return pm_ref


def _conf_file(tmp_path, text):
    p = tmp_path / "particle_mesh.conf"
    p.write_text(text, encoding="utf-8")
    return str(p)


def _outputs(outdir):
    files = []
    for dp, _, fs in os.walk(outdir):
        files += [os.path.join(dp, f) for f in fs if f.endswith(".npz")]
    return sorted(files)


@pytest.mark.parametrize("device_ics", [False, True], ids=["host_ics", "device_ics"])
def test_run_noname_shipped_cosmology_conf(lib_built, O, tmp_path, device_ics):
    """BASELINE config 1b: run/CosmologyPM as shipped (2-D, 128^2 particles on a 64^2 mesh, EdS Zel'dovich pancake,
    z = 400 -> 9, ~7600 steps of evolvePMCosmology's dt control) end to end against the oracle's loop."""
    from noname_b200 import run_noname
    outdir = str(tmp_path / "out")
    over = {"OutputDirectory": outdir, "RestartFileName": str(tmp_path / "restart.bin"), "SortEvery": 8,
            "OutputFormat": "npz"}
    if device_ics:
        over["DeviceInitialConditions"] = True
    gp, gd, p = run_noname(_conf_file(tmp_path, SHIPPED_COSMOLOGY_PM), overrides=over)
    from noname_b200 import NoName
    c = NoName.conf

    oc = O.Cosmo(h=1.0, OmegaM=1.0, OmegaR=0.0)
    ts, te = O.start_stop_times(oc, 400, 9)
    a = O.a_from_t_internal(oc, ts, 400, zwherea1=400)
    ad = O.dadt_from_t_internal(oc, ts, 400)
    xr, vr = O.zeldovich_pancake((128, 128, 1), 2, 10.0, 400, a, ad)
    m = 64 * 64 / (128.0 * 128.0)
    rc = dict(StartTime=ts, StopTime=te, StartCycle=0, StopCycle=1000000, InitialDt=1e-4, CourantFactor=0.5, MaximumDt=1.0,
              DtFractionOfTime=0.1, InitialRedshift=400)
    t, cyc, f = O.evolvePMCosmology(xr, vr, m, (64, 64, 1), rc, oc)
    assert c["CurrentCycle"] == cyc and abs(c["CurrentTime"] - t) <= 1e-9 * t
    assert p["m"] == pytest.approx(m)
    ex, ev = float(np.max(np.abs(p["x"] - xr))), nlinf(p["v"], vr)
    ephi, erho = nlinf(gd[1].d["Φ"], f.phi), nlinf(gd[1].d["ρD"], f.rho)
    print(f"shipped CosmologyPM conf: {cyc} steps, |dx| {ex:.2e} dv {ev:.2e} dphi {ephi:.2e} drho {erho:.2e}")
    # ~7600 steps through a caustic (zc = 10 > z_final = 9); measured on the B200: |dx| 2e-14, dv 2e-13, dphi 2e-14, drho 1e-13
    assert ex < 1e-11 and ev < 1e-10
    assert ephi < 1e-10 and erho < 1e-10
    assert gd[1].d["ρD"].sum() == pytest.approx(m * 128 * 128, rel=1e-10)       # the real density, not the initial ones()
    # P(k) of the final state within 1e-6 (north star)
    from noname_b200 import ParticleMesh as PM
    kg, pg = PM.powerSpectrum(gp, gd, p, c)
    kr, pr = O.powerSpectrum(xr.copy(), m, (64, 64, 1), "cic")
    ok = np.isfinite(pr)
    assert np.max(np.abs(pg[ok] / pr[ok] - 1.0)) < 1e-6
    # outputs: OutputDtDataDump = 3.5 -> many dumps, each holding the density of the step it follows
    outs = _outputs(outdir)
    assert len(outs) >= 3
    last = np.load(outs[-1])
    assert last["grid1/ρD"].sum() == pytest.approx(m * 128 * 128, rel=1e-10)
    assert np.array_equal(last["grid1/ρD"], np.squeeze(gd[1].d["ρD"]))            # the forced final dump
    first = np.load(outs[0])
    assert first["grid1/ρD"].sum() == pytest.approx(m * 128 * 128, rel=1e-10) and not np.allclose(first["grid1/ρD"], 1.0)
    assert os.path.exists(over["RestartFileName"])


def test_run_noname_shipped_pm_conf_shape(lib_built, O, tmp_path):
    """BASELINE config 1: run/PM as shipped -- static evolvePM, 100^2 particles on a 1000 x 1000 x 1 mesh (non-power-of-two:
    cuFFT 2-D + Green kernel fallback, every tile box wraps), CourantFactor 0.2, MaximumDt 0.1, StopTime 1, dumps every
    0.5.  The shipped ProblemDefinition does not parse in the reference (ParticleMesh.jl:246) and names multiplePeaksPS;
    it runs here with InputPowerSpectrum = scaleFreePS(-1, .3) (commented out in the shipped CosmologyPM conf) through the
    device generator, whose lattice (100 x 100) differs from the mesh: the temporary lattice context is exercised."""
    from noname_b200 import run_noname, NoName
    outdir = str(tmp_path / "out")
    over = {"OutputDirectory": outdir, "RestartFileName": str(tmp_path / "restart.bin"), "InputPowerSpectrum": "scaleFreePS(-1, .3)",
            "OutputFormat": "npz"}
    gp, gd, p = run_noname(_conf_file(tmp_path, SHIPPED_PM), overrides=over)
    c = NoName.conf
    xr, vr = O.synthetic_grf((100, 100, 1), 2, 12345, -1.0, 0.3 / 1000, 0.0)
    m = 1000.0 * 1000.0 / 1e4
    rc = dict(StartTime=0.0, StopTime=1.0, StartCycle=0, StopCycle=1000000, InitialDt=1e-7, CourantFactor=0.2, MaximumDt=0.1)
    t, cyc, f = O.evolvePM(xr, vr, m, (1000, 1000, 1), rc)
    assert c["CurrentCycle"] == cyc == 11 and abs(c["CurrentTime"] - t) <= 1e-12
    assert p["x"].shape == (10000, 2) and p["m"] == m
    assert float(np.max(np.abs(p["x"] - xr))) < 1e-11 and nlinf(p["v"], vr) < 1e-9
    assert nlinf(gd[1].d["Φ"], f.phi) < 1e-10 and nlinf(gd[1].d["ρD"], f.rho) < 1e-10
    outs = _outputs(outdir)
    # OutputDtDataDump = 0.5 against LastOutputTime = 0 (default.conf): the dump falls at t = 0.5 + InitialDt; the last step
    # lands on StopTime = 1.0, 0.4999999 after it -- no second dump, and evolvePM has no forced final one (:780-796)
    assert len(outs) == 1
    for o in outs:
        assert np.load(o)["grid1/ρD"].sum() == pytest.approx(m * 1e4, rel=1e-10)


@pytest.mark.parametrize("pdims,dims,rank", [((16, 16, 16), (16, 16, 16), 3), ((128, 64, 64), (128, 64, 64), 3),
                                             ((16, 16, 16), (32, 32, 32), 3), ((24, 20, 1), (16, 16, 1), 2),
                                             ((128, 64, 1), (128, 64, 1), 2)])
def test_grf_initial_conditions_match_oracle(lib_built, O, pdims, dims, rank):
    """nnpm_init_grf (PowerSpectrumParticles + scaleFreePS, seeded) == oracle twin: library FFT fallbacks (16^3), the own
    x / y passes (128 x 64 x 64), a particle lattice that differs from the mesh (temporary lattice context), 2-D."""
    from noname_b200 import DevicePM
    npart = int(np.prod(pdims))
    xr, vr = O.synthetic_grf(pdims, rank, 4242, -1.0, 0.3 / max(dims), 0.7)
    with DevicePM(dims, npart, rank=rank) as dev:
        dev.init_grf(pdims, seed=4242, ps_index=-1.0, disp_rms=0.3 / max(dims), vfac=0.7, m=1.0)
        xg, vg = dev.download()
        d = np.abs(xg - xr)
        d = np.minimum(d, 1.0 - d)                         # a particle within rounding of the periodic edge may wrap
        assert float(d.max()) < 1e-13
        assert nlinf(vg, vr) < 1e-11
        rms = np.sqrt(np.mean((vg / 0.7) ** 2))
        assert rms == pytest.approx(0.3 / max(dims), rel=1e-12)
        st = dev.step_comoving(1e-3, 1.0, 1.0, 0.5, 1.0)      # the generated state is steppable (mesh left clean)
        assert st.n_halo_violations == 0 and np.isfinite(st.max_abs_v)


PM_PEAKS = ([0.1, 1.0, 10.0], [1.0, 1.0, 1.0])      # multiplePeaksPS(-1, 1, [0.1, 1, 10], 1.) of run/PM/particle_mesh.conf


@pytest.mark.parametrize("pdims,dims,rank", [((100, 100, 1), (1000, 1000, 1), 2), ((128, 64, 64), (128, 64, 64), 3),
                                             ((32, 32, 32), (32, 32, 32), 3)])
def test_grf_multiple_peaks_initial_conditions_match_oracle(lib_built, O, pdims, dims, rank):
    """nnpm_init_grf_peaks (PowerSpectrumParticles + multiplePeaksPS, src/InputPowerSpectra.jl:6-36) == oracle twin: the
    run/PM shape (100^2 lattice on a 1000^2 mesh: library transforms on a temporary lattice context), the own x / y / z
    passes, and a small cube; a spectrum with no lattice mode inside its windows leaves the lattice untouched."""
    from noname_b200 import DevicePM
    npart = int(np.prod(pdims))
    xr, vr = O.synthetic_grf(pdims, rank, 777, -1.0, 0.3 / max(dims), 0.7, peaks=PM_PEAKS)
    xf, _ = O.synthetic_grf(pdims, rank, 777, -1.0, 0.3 / max(dims), 0.7)
    assert float(np.max(np.abs(xr - xf))) > 1e-6            # the windows do change the field
    with DevicePM(dims, npart, rank=rank) as dev:
        dev.init_grf(pdims, seed=777, ps_index=-1.0, disp_rms=0.3 / max(dims), vfac=0.7, m=1.0, peaks=PM_PEAKS)
        xg, vg = dev.download()
        d = np.abs(xg - xr)
        d = np.minimum(d, 1.0 - d)
        assert float(d.max()) < 1e-13
        assert nlinf(vg, vr) < 1e-11
        dev.init_grf(pdims, seed=777, ps_index=-1.0, disp_rms=0.3 / max(dims), vfac=0.7, m=1.0, peaks=([50.0], [0.01]))
        x0, v0 = dev.download()
        assert np.array_equal(x0, O.initialize_particles_uniform(pdims, rank)) and not v0.any()
        with pytest.raises(Exception):
            dev.init_grf(pdims, peaks=([1.0] * 9, [1.0] * 9))      # more than NNPM_MAX_PEAKS


def test_run_noname_shipped_pm_conf_verbatim(lib_built, O, tmp_path):
    """BASELINE config 1 with NOTHING overridden but the output paths: run/PM/particle_mesh.conf as shipped, including its
    InputPowerSpectrum = multiplePeaksPS(-1, 1, [0.1, 1, 10], 1.), end to end against the oracle's evolvePM."""
    from noname_b200 import run_noname, NoName
    over = {"OutputDirectory": str(tmp_path / "out"), "RestartFileName": str(tmp_path / "restart.bin"), "OutputFormat": "npz"}
    gp, gd, p = run_noname(_conf_file(tmp_path, SHIPPED_PM), overrides=over)
    c = NoName.conf
    xr, vr = O.synthetic_grf((100, 100, 1), 2, 12345, -1.0, 0.3 / 1000, 0.0, peaks=PM_PEAKS)
    m = 1000.0 * 1000.0 / 1e4
    rc = dict(StartTime=0.0, StopTime=1.0, StartCycle=0, StopCycle=1000000, InitialDt=1e-7, CourantFactor=0.2, MaximumDt=0.1)
    t, cyc, f = O.evolvePM(xr, vr, m, (1000, 1000, 1), rc)
    assert c["CurrentCycle"] == cyc and abs(c["CurrentTime"] - t) <= 1e-12
    assert float(np.max(np.abs(p["x"] - xr))) < 1e-11 and nlinf(p["v"], vr) < 1e-9
    assert nlinf(gd[1].d["Φ"], f.phi) < 1e-10 and nlinf(gd[1].d["ρD"], f.rho) < 1e-10
    assert len(_outputs(str(tmp_path / "out"))) >= 1


def test_pancake_device_matches_oracle(lib_built, O):
    from noname_b200 import DevicePM
    oc = O.Cosmo(h=1.0, OmegaM=1.0, OmegaR=0.0)
    ts, _ = O.start_stop_times(oc, 400, 9)
    a = O.a_from_t_internal(oc, ts, 400, zwherea1=400)
    ad = O.dadt_from_t_internal(oc, ts, 400)
    A = 5.0 * 11.0 / 2.0 / 401.0
    for pdims, dims, rank in (((128, 128, 1), (64, 64, 1), 2), ((16, 8, 8), (16, 16, 16), 3)):
        xr, vr = O.zeldovich_pancake(pdims, rank, 10.0, 400, a, ad)
        with DevicePM(dims, int(np.prod(pdims)), rank=rank) as dev:
            dev.init_pancake(pdims, 0.4 * A * a, 0.4 * A * ad, m=1.0)
            xg, vg = dev.download()
        assert float(np.max(np.abs(xg - xr))) < 1e-15 and float(np.max(np.abs(vg - vr))) < 1e-15 * max(1.0, np.abs(vr).max())


def test_single_step_fields_at_512_cubed_against_threaded_c_port(lib_built):
    """north star: single-step density and acceleration fields within 1e-12 relative -- checked at a BASELINE size
    (config 3: 512^3 particles on a 512^3 mesh) against the C restatement run on the host cores.  The particles are
    the device-generated random-field ICs, downloaded so that both sides start from identical arrays."""
    import gc
    from noname_b200 import DevicePM
    from oracle.pm_ref_c import PMRefC
    n = 512
    threads = max(1, min(32, len(os.sched_getaffinity(0))))
    ref = PMRefC(threads=threads)
    with DevicePM((n, n, n), n ** 3, rank=3, sort_every=1) as dev:
        dev.init_grf((n, n, n), seed=12345, ps_index=-1.0, disp_rms=0.3 / n, vfac=0.0, m=0.3)
        x, v = dev.download()
        dev.sort()                                   # the tiled deposit is the path under test
        dev.deposit()
        rho = dev.get_field("rho")
        dev.solve()
        phi = dev.get_field("phi")
        dev.accel()
        acc = dev.get_field("acc")
    del v
    gc.collect()
    shape = (n, n, n)
    rrho = np.zeros(shape)
    ref.make_periodic(x)
    ref.deposit(rrho, x, 0.3, "cic")
    e_rho = nlinf(rho, rrho)
    assert abs(rho.sum() - 0.3 * n ** 3) < 1e-10 * 0.3 * n ** 3
    del x, rho
    tmp = np.empty(shape)
    ref.sub_mean(tmp, rrho)
    del rrho
    rt = np.zeros((n, n, n // 2 + 1), dtype=np.complex128)
    rphi = np.empty(shape)
    ref.compute_potential(rphi, tmp, rt)
    del tmp, rt
    e_phi = nlinf(phi, rphi)
    del phi
    racc = np.zeros(shape + (3,))
    ref.compute_acceleration(racc, rphi)
    e_acc = nlinf(acc, racc)
    print(f"512^3 single step vs C port ({threads} threads): rho {e_rho:.2e} phi {e_phi:.2e} acc {e_acc:.2e}")
    assert e_rho < 1e-12 and e_phi < 1e-12 and e_acc < 1e-12
