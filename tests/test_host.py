"""Host-side logic that needs no GPU: conf parsing / layering / aliases (src/NoName.jl:58-72,
src/default.conf), cosmology scalars (src/cosmology.jl) against the oracle's independent
quad+brentq restatement, output cadence (src/IOtools.jl:40-59), initialisers (ParticleMesh.jl:103-287)
and the evolve loops' time-step control driven through a recording stand-in for the device."""
import math
import os

import numpy as np
import pytest

from noname_b200 import conf as cf
from noname_b200 import cosmology as cos
from noname_b200 import ParticleMesh as PM
from noname_b200 import NoName
from oracle import pm_ref as O

SHIPPED_COSMO_CONF = """
verbose = false
Initializer = InitializeParticleSimulation # function for Problem initialization
ProblemDefinition = ZeldovichPancakeParticles # string identifier
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
ParticleBackInterpolation    = cic
OutputDirPrefix   = "pm_"
OutputDtDataDump = 3.5
Cosmology = cosmology(h=1.,OmegaM=1.0,OmegaR=0.0) # see Cosmology.jl
OmegaBaryon = 0.
ComovingCoordinates = true
InitialRedshift = 400
FinalRedshift   = 9
"""


def _write(tmp_path, text, name="c.conf"):
    p = tmp_path / name
    p.write_text(text, encoding="utf-8")
    return str(p)


# ------------------------------------------------------------------------------ conf
def test_parse_values():
    assert cf.parse_value("[64, 64, 1]") == [64, 64, 1]
    assert cf.parse_value("0.0000001") == 1e-7
    assert cf.parse_value("1e30") == 1e30 and cf.parse_value("10.") == 10.0
    assert cf.parse_value("true") is True and cf.parse_value("false") is False
    assert cf.parse_value('"pm_"') == "pm_"
    assert cf.parse_value("evolvePM") == "evolvePM"
    assert cf.parse_value("cosmology(h=1.,OmegaM=1.0,OmegaR=0.0)") == "cosmology(h=1.,OmegaM=1.0,OmegaR=0.0)"
    assert cf.parse_value("1_000") == 1000


def test_conf_layering_and_comments(tmp_path):
    c = cf.load_conf(_write(tmp_path, SHIPPED_COSMO_CONF), user_default=False)
    assert c["TopgridDimensions"] == [64, 64, 1] and c["ParticleDimensions"] == [128, 128, 1]
    assert c["Evolve"] == "evolvePMCosmology" and c["CourantFactor"] == 0.5
    assert c["ParticleDepositInterpolation"] == "cic" and c["OutputDirPrefix"] == "pm_"
    assert c["StopCycle"] == 1000000 and c["RestartFileName"] == "/tmp/restart.jld"   # defaults kept
    assert c["InitialRedshift"] == 400 and c["DtFractionOfTime"] == 0.1
    c2 = cf.load_conf(_write(tmp_path, SHIPPED_COSMO_CONF), overrides={"CourantFactor": 0.25}, user_default=False)
    assert c2["CourantFactor"] == 0.25


def test_conf_aliases_named_by_the_north_star(tmp_path):
    c = cf.load_conf(_write(tmp_path, "dims = [32, 32, 32]\ncosm_Ωm = 0.3\ncosm_h = 0.7\n"), user_default=False)
    assert c["TopgridDimensions"] == [32, 32, 32]
    co = cos.parse_cosmology(c["Cosmology"])
    assert (co.Om, co.h, co.Or) == (0.3, 0.7, 0.0) and co.OL == pytest.approx(0.7)


def test_reference_defaults_verbatim():
    d = cf.DEFAULTS
    assert d["CourantFactor"] == 0.3 and d["InitialDt"] == 1e-7 and d["MaximumDt"] == 1e30
    assert d["ParticleDepositInterpolation"] == "cic" and d["ParticleBackInterpolation"] == "cic"
    assert d["DtFractionOfTime"] == 10000.0 and d["OutputDirectory"] == "/tmp/noname"


# ------------------------------------------------------------------------------ cosmology
@pytest.mark.parametrize("kw", [dict(h=1.0, OmegaM=1.0, OmegaR=0.0), dict(h=0.7, OmegaM=0.3, OmegaR=0.0),
                                dict(h=0.69, OmegaM=0.29), dict(h=0.7, OmegaM=0.3, OmegaK=0.05, OmegaR=1e-4)])
def test_cosmology_scalars_match_oracle(kw):
    c, o = cos.Cosmology(**kw), O.Cosmo(**kw)
    for z in (400.0, 50.0, 9.0, 1.0, 0.0):
        assert c.age_gyr(z) == pytest.approx(o.age_gyr(z), rel=2e-12)
    zi = 50.0
    ts, te = O.start_stop_times(o, zi, 0.0)
    for t in (ts, 1.7 * ts, 0.5 * (ts + te), te):
        assert cos.a_from_t_internal(c, t, zi, zwherea1=zi) == pytest.approx(O.a_from_t_internal(o, t, zi, zwherea1=zi), rel=1e-11)
        assert cos.dadt_from_t_internal(c, t, zi) == pytest.approx(O.dadt_from_t_internal(o, t, zi), rel=1e-11)
        assert cos.z_from_t_internal(c, t, zi) == pytest.approx(O.z_from_t_internal(o, t, zi), rel=1e-10, abs=1e-11)


def test_cosmology_eds_closed_form():
    c = cos.Cosmology(h=1.0, OmegaM=1.0, OmegaR=0.0)
    zi = 400.0
    u = cos.get_units(c, zi, zinit=zi)
    ts = c.age_gyr(zi) * (cos.gyr_in_s / u.T)
    assert ts == pytest.approx(math.sqrt(2.0 / 3.0), rel=5e-6)           # cosmology.jl:69 constants pin t_H (test_kat8)
    assert cos.HUBBLE_TIME_GYR == O.HUBBLE_TIME_GYR
    old = cos.Cosmology(h=1.0, OmegaM=1.0, OmegaR=0.0, hubble_time_gyr=9.77793)   # rounds 1-2 (bench.py's stored P(k))
    assert old.age_gyr(zi) / c.age_gyr(zi) == pytest.approx(9.77793 / 9.77813952, rel=1e-15)
    assert cos.a_from_t_internal(c, 27 * ts, zi, zwherea1=zi) == pytest.approx(9.0, rel=1e-12)
    assert cos.dadt(c, 4.0, zwherea1=zi) == math.sqrt(2.0 / (3.0 * 4.0))
    assert cos.z_from_a(c, 2.0, zwherea1=3.0) == 1.0


def test_parse_cosmology_rejects_garbage():
    with pytest.raises(ValueError):
        cos.parse_cosmology("notacosmology(1)")
    with pytest.raises(ValueError):
        cos.parse_cosmology("cosmology(w0=-1)")


# ------------------------------------------------------------------------------ output cadence
def test_ready_for_output_rules():
    c = dict(OutputEveryNCycle=0, OutputDtDataDump=0.0, CurrentCycle=5, CurrentTime=1.0, LastOutputTime=0.0)
    assert PM.readyForOutput(c) is False
    c["OutputEveryNCycle"] = 5
    assert PM.readyForOutput(c) is True and c["LastOutputCycle"] == 5 and c["LastOutputTime"] == 1.0
    c.update(OutputEveryNCycle=0, OutputDtDataDump=0.5, CurrentTime=1.4)
    assert PM.readyForOutput(c) is False
    c["CurrentTime"] = 1.5
    assert PM.readyForOutput(c) is True


# ------------------------------------------------------------------------------ initialisers
def test_initialize_particle_simulation_matches_reference_rules(tmp_path):
    c = cf.load_conf(_write(tmp_path, SHIPPED_COSMO_CONF), user_default=False)
    gp, gd, p = {}, {}, {}
    PM.InitializeParticleSimulation(gp, gd, p, c)
    assert c["_rank"] == 2 and gp[1].dim == (64, 64, 1)
    assert gd[1].d["ρD"].shape == (1, 64, 64) and gd[1].d["Φ"].shape == (1, 64, 64)
    assert p["x"].shape == (128 * 128, 2)
    assert p["m"] == pytest.approx(64 * 64 / (128.0 * 128.0) * 1.0)       # Ncell/Npart * OmegaCDM
    assert c["StartTime"] == pytest.approx(math.sqrt(2.0 / 3.0), rel=2e-4)
    assert c["StopTime"] > c["StartTime"]
    # pancake ICs equal the oracle's restatement of :177-198
    oc = O.Cosmo(h=1.0, OmegaM=1.0, OmegaR=0.0)
    a = O.a_from_t_internal(oc, c["StartTime"], 400, zwherea1=400)
    ad = O.dadt_from_t_internal(oc, c["StartTime"], 400)
    xr, vr = O.zeldovich_pancake((128, 128, 1), 2, 10.0, 400, a, ad)
    assert np.max(np.abs(p["x"] - xr)) < 1e-12 and np.max(np.abs(p["v"] - vr)) < 1e-12


def test_lattice_matches_oracle():
    x = np.zeros((4 * 3 * 2, 3))
    PM.initialize_particles_uniform(x, [4, 3, 2])
    assert np.array_equal(x, O.initialize_particles_uniform((4, 3, 2), 3))
    assert x[0].tolist() == [0.125, 1.0 / 6.0, 0.25] and x[1, 0] == 0.375   # i fastest


def test_unknown_or_broken_problem_definitions_raise(tmp_path):
    c = cf.load_conf(_write(tmp_path, "TopgridDimensions=[8,8,1]\nParticleDimensions=[8,8,1]\nProblemDefinition = PowerSpectrumParticles\n"), user_default=False)
    with pytest.raises(KeyError):                 # :212-216: InputPowerSpectrum is required
        PM.InitializeParticleSimulation({}, {}, {}, c)
    c["InputPowerSpectrum"] = "tophatPS(1, 2)"                               # not a spectrum of InputPowerSpectra.jl
    with pytest.raises(NotImplementedError):
        PM.InitializeParticleSimulation({}, {}, {}, c)
    c["InputPowerSpectrum"] = "multiplePeaksPS(-1, 1, [0.1, 1, 10], [1., 2.])"   # widths must match the peaks (:20-24)
    with pytest.raises(ValueError):
        PM.InitializeParticleSimulation({}, {}, {}, c)
    c["InputPowerSpectrum"] = "scaleFreePS(-1, .3)"
    c["ProblemDefinition"] = "Nope"
    with pytest.raises(KeyError):
        PM.InitializeParticleSimulation({}, {}, {}, c)
    with pytest.raises(NotImplementedError):
        PM.deposit(np.zeros((1, 4, 4)), np.zeros((1, 2)), 1.0, interpolation="sic")
    with pytest.raises(ValueError):
        PM.deposit(np.zeros((1, 4, 4)), np.zeros((1, 2)), 1.0, interpolation="tsc")


def test_power_spectrum_particles_are_generated_on_the_device(tmp_path):
    """PowerSpectrumParticles + scaleFreePS (ParticleMesh.jl:209-254, InputPowerSpectra.jl:26): the host only records
    the request (seed, spectral index, RMS displacement, growing-mode velocity factor); nothing is allocated"""
    c = cf.load_conf(_write(tmp_path, SHIPPED_COSMO_CONF.replace("ProblemDefinition = ZeldovichPancakeParticles",
                                                                  "ProblemDefinition = PowerSpectrumParticles\nInputPowerSpectrum = scaleFreePS(-1, .3)")),
                     user_default=False)
    gp, gd, p = {}, {}, {}
    PM.InitializeParticleSimulation(gp, gd, p, c)
    assert p["x"] is None and p["v"] is None
    kind, kw = p["_ic"]
    oc = O.Cosmo(h=1.0, OmegaM=1.0, OmegaR=0.0)
    a = O.a_from_t_internal(oc, c["StartTime"], 400, zwherea1=400)
    ad = O.dadt_from_t_internal(oc, c["StartTime"], 400)
    assert kind == "grf" and kw["ps_index"] == -1.0 and kw["seed"] == 12345
    assert kw["disp_rms"] == pytest.approx(0.3 / 64) and kw["vfac"] == pytest.approx(ad / a, rel=1e-9)


def test_multiple_peaks_spectrum_as_shipped_in_run_pm(tmp_path):
    """run/PM/particle_mesh.conf names InputPowerSpectrum = multiplePeaksPS(-1, 1, [0.1, 1, 10], 1.) (src/InputPowerSpectra.jl:
    6-24): parsed into (n, norm, peaks); a scalar width applies to every peak; the request goes to nnpm_init_grf_peaks."""
    assert PM._parse_input_power_spectrum("multiplePeaksPS(-1, 1, [0.1, 1, 10], 1.)") == (-1.0, 1.0, ([0.1, 1.0, 10.0], [1.0, 1.0, 1.0]))
    assert PM._parse_input_power_spectrum(" multiplePeaksPS(-2,3.5,[1, 10, 30],[0.1, 0.2, 0.3]) ") == (-2.0, 3.5, ([1.0, 10.0, 30.0], [0.1, 0.2, 0.3]))
    assert PM._parse_input_power_spectrum("scaleFreePS(-1, .3)") == (-1.0, 0.3, None)
    c = cf.load_conf(_write(tmp_path, "TopgridDimensions=[16,16,1]\nParticleDimensions=[8,8,1]\nProblemDefinition = PowerSpectrumParticles\n"
                                      "InputPowerSpectrum = multiplePeaksPS(-1, 1, [0.1, 1, 10], 1.)\n"), user_default=False)
    gp, gd, p = {}, {}, {}
    PM.InitializeParticleSimulation(gp, gd, p, c)
    assert p["x"] is None and p["_ic"][0] == "grf" and p["_ic"][1]["peaks"] == ([0.1, 1.0, 10.0], [1.0, 1.0, 1.0])
    assert p["_ic"][1]["vfac"] == 0.0 and p["_ic"][1]["disp_rms"] == pytest.approx(0.3 / 16)


def test_oracle_multiple_peaks_spectrum_is_zero_outside_the_windows():
    """checker of nnpm_init_grf_peaks: P = k^n where |k - k_p| / k < w_p / 2 (InputPowerSpectra.jl:28-36), else 0 -- and
    identical to the scale-free field wherever it is non-zero (same seeded draws per mode)"""
    pd = (32, 24, 1)
    peaks = ([0.1, 1.0, 10.0], [1.0, 1.0, 1.0])
    free = O.grf_spectrum(pd, 0, 99, -1.0)
    pk = O.grf_spectrum(pd, 0, 99, -1.0, peaks)
    K, J, I = np.meshgrid(np.arange(1), np.arange(24), np.arange(17), indexing="ij")
    kx = 2 * math.pi * I / 32
    ky = 2 * math.pi * np.where(J > 12, J - 24, J) / 24
    ka = np.sqrt(kx * kx + ky * ky)
    with np.errstate(invalid="ignore", divide="ignore"):
        inside = ((np.abs(ka - 0.1) / ka < 0.5) | (np.abs(ka - 1.0) / ka < 0.5) | (np.abs(ka - 10.0) / ka < 0.5)) & (ka > 0)
    assert inside.any() and (~inside).any()
    assert np.all(pk[~inside] == 0) and np.array_equal(pk[inside], free[inside])
    x, v = O.synthetic_grf(pd, 2, 99, -1.0, 0.01, 0.5, peaks=peaks)
    q = O.initialize_particles_uniform(pd, 2)
    d = x - q
    d -= np.rint(d)
    assert np.sqrt(np.mean(d * d)) == pytest.approx(0.01, rel=1e-12) and np.allclose(v, 0.5 * d, atol=1e-15)
    x0, _ = O.synthetic_grf(pd, 2, 99, -1.0, 0.01, 0.5, peaks=([10.0], [0.1]))      # no lattice mode inside: zero displacement
    assert np.array_equal(x0, q)


def test_device_initial_conditions_key_defers_pancake_and_lattice(tmp_path):
    c = cf.load_conf(_write(tmp_path, SHIPPED_COSMO_CONF + "DeviceInitialConditions = true\n"), user_default=False)
    gp, gd, p = {}, {}, {}
    PM.InitializeParticleSimulation(gp, gd, p, c)
    assert p["x"] is None and p["_ic"][0] == "pancake"
    oc = O.Cosmo(h=1.0, OmegaM=1.0, OmegaR=0.0)
    A = 5.0 * 11.0 / 2.0 / 401.0
    assert p["_ic"][1]["xamp"] == pytest.approx(0.4 * A * O.a_from_t_internal(oc, c["StartTime"], 400, zwherea1=400), rel=1e-9)
    assert p["_ic"][1]["vamp"] == pytest.approx(0.4 * A * O.dadt_from_t_internal(oc, c["StartTime"], 400), rel=1e-9)


# ------------------------------------------------------------------------------ evolve loops: dt control
class _FakeStats:
    def __init__(self, mv):
        self.max_abs_v, self.max_v, self.max_pa = mv, mv, 0.0


class _FakeDev:
    """records the scalars the loops hand to the device; returns scripted max|v| values"""
    rank = 2
    nranks = 1
    rank_id = 0

    def __init__(self, maxv):
        self.maxv, self.calls, self.keep = maxv, [], []

    def set_option(self, name, value):
        assert name == "keep_rho"
        self.keep.append(bool(value))

    def step_static(self, dt):
        self.calls.append((dt,))
        return _FakeStats(self.maxv)

    def step_comoving(self, dt, a1, ak, adk, a2):
        self.calls.append((dt, a1, ak, adk, a2))
        return _FakeStats(self.maxv)

    local_count = 0

    def download(self, x, v):
        pass

    def get_field(self, *a):
        pass

    def close(self):
        pass


def test_evolvepm_dt_control_matches_oracle_rules(monkeypatch):
    """ParticleMesh.jl:780-789: Courant, MaximumDt, final-step clamp (1+1e-15)."""
    fake = _FakeDev(2.0)
    monkeypatch.setattr(PM, "_device_for", lambda gp, p, c: fake)
    monkeypatch.setattr(PM, "_sync_host", lambda *a, **k: None)
    c = dict(cf.DEFAULTS, StartTime=0.0, StopTime=0.1, InitialDt=1e-3, CourantFactor=0.5, MaximumDt=0.03,
             OutputDisabled=True, ParticleDimensions=[8, 8, 1])
    gp = {1: PM.GridPatch([0, 0, 0], [1, 1, 1], 1, 1, (16, 16, 1))}
    PM.evolvePM(gp, {}, {}, c)
    dts = [k[0] for k in fake.calls]
    courant = 0.5 * (1.0 / 16) / (2.0 + 1e-30)
    assert dts[0] == 1e-3 and all(d == pytest.approx(courant) for d in dts[1:-1])
    assert sum(dts) == pytest.approx(0.1, rel=1e-12) and c["CurrentTime"] >= 0.1
    assert dts[-1] <= courant * (1 + 1e-12)
    assert c["CurrentCycle"] == len(dts)
    # outputs disabled: the density is kept (copied on the device) for the last step only
    assert fake.keep == [False] * (len(dts) - 1) + [True]


def test_keep_rho_is_requested_exactly_for_steps_followed_by_an_output(monkeypatch, tmp_path):
    """gd[1].d["ρD"] of the reference holds the last deposit whenever analyzePM writes (ParticleMesh.jl:767 / :857,
    IOtools.jl:94-116): the loop asks the device to keep the density for those steps and no others"""
    fake = _FakeDev(2.0)
    monkeypatch.setattr(PM, "_device_for", lambda gp, p, c: fake)
    synced = []
    monkeypatch.setattr(PM, "_sync_host", lambda *a, **k: synced.append(len(fake.calls)))
    c = dict(cf.DEFAULTS, StartTime=0.0, StopTime=1.0, StopCycle=7, InitialDt=1e-3, CourantFactor=0.5, MaximumDt=0.03,
             OutputEveryNCycle=3, OutputDirectory=str(tmp_path / "out"), ParticleDimensions=[8, 8, 1])
    gp = {1: PM.GridPatch([0, 0, 0], [1, 1, 1], 1, 1, (16, 16, 1))}
    gd = {1: PM.GridData({"ρD": np.ones((1, 16, 16)), "Φ": np.ones((1, 16, 16))})}
    p = {"x": np.zeros((4, 2)), "v": np.zeros((4, 2)), "m": 1.0}
    PM.evolvePM(gp, gd, p, c)
    assert len(fake.calls) == 7
    assert fake.keep == [False, False, True, False, False, True, True]     # cycles 3 and 6 write; 7 ends the loop
    assert synced == [3, 6, 7]


def test_evolvepmcosmology_scalars_and_dt(monkeypatch, tmp_path):
    """:853-885: a(t+dt/4), a,adot(t+dt/2), a(t+3dt/4); dt = min(Courant, frac*t, MaximumDt); (1+1e-6) clamp."""
    fake = _FakeDev(0.05)
    monkeypatch.setattr(PM, "_device_for", lambda gp, p, c: fake)
    monkeypatch.setattr(PM, "_sync_host", lambda *a, **k: None)
    c = cf.load_conf(_write(tmp_path, SHIPPED_COSMO_CONF), user_default=False)
    c["OutputDisabled"] = True
    c["StopCycle"] = 5
    gp, gd, p = {}, {}, {}
    PM.InitializeParticleSimulation(gp, gd, p, c)
    PM.evolvePMCosmology(gp, gd, p, c)
    oc = O.Cosmo(h=1.0, OmegaM=1.0, OmegaR=0.0)
    t, dt = c["StartTime"], 1e-4
    assert len(fake.calls) == 5
    for (gdt, a1, ak, adk, a2) in fake.calls:
        assert gdt == pytest.approx(dt, rel=1e-13)
        assert a1 == pytest.approx(O.a_from_t_internal(oc, t + dt / 4, 400, zwherea1=400), rel=1e-10)
        assert ak == pytest.approx(O.a_from_t_internal(oc, t + dt / 2, 400, zwherea1=400), rel=1e-10)
        assert adk == pytest.approx(O.dadt_from_t_internal(oc, t + dt / 2, 400), rel=1e-10)
        assert a2 == pytest.approx(O.a_from_t_internal(oc, t + 0.75 * dt, 400, zwherea1=400), rel=1e-10)
        t += dt
        dt = min(0.5 * (1.0 / 64) / (0.05 + 1e-30), 0.1 * t, 1.0)


def test_analyzepm_writes_reference_dataset_names_and_power_spectrum(monkeypatch, tmp_path):
    """:13-63 + IOtools.jl:73-116: NNNNN_parameters.info, the data file under the reference's dataset names, and -- with
    OutputPowerSpectrum = true -- the P(k) numbers the reference would have plotted (:47-61)."""
    class Dev(_FakeDev):
        def power_spectrum(self):
            self.calls.append("pk")
            return np.array([1.0, 2.0]), np.array([0.5, 0.25])
    dev = Dev(1.0)
    synced = []
    monkeypatch.setattr(PM, "_sync_host", lambda *a, **k: synced.append(len(dev.calls)))
    c = dict(cf.DEFAULTS, OutputDirectory=str(tmp_path / "out"), OutputDirPrefix="pm_", OutputPowerSpectrum=True,
             OutputFormat="npz", CurrentTime=0.75, _hidden=1)
    gp = {1: PM.GridPatch([0, 0, 0], [1, 1, 1], 1, 1, (4, 4, 1))}
    gd = {1: PM.GridData({"ρD": np.full((1, 4, 4), 2.0), "Φ": np.zeros((1, 4, 4))})}
    p = {"x": np.arange(6.0).reshape(3, 2), "v": -np.arange(6.0).reshape(3, 2), "m": 0.5}
    assert PM.analyzePM(gp, gd, p, c, dev) is None and not synced          # nothing due, nothing copied
    PM.analyzePM(gp, gd, p, c, dev, force=True)
    assert synced == [0] and dev.calls == ["pk"]                            # host state first, then the P(k) deposit
    d = tmp_path / "out" / "pm_00000"
    assert sorted(f.name for f in d.iterdir()) == ["00000.npz", "00000_PS.txt", "00000_parameters.info"]
    z = np.load(d / "00000.npz")
    assert sorted(z.files) == ["grid1/Φ", "grid1/ρD", "particles/m", "particles/vx", "particles/vy", "particles/x", "particles/y"]
    assert z["grid1/ρD"].shape == (4, 4) and z["particles/y"].tolist() == [1.0, 3.0, 5.0] and float(z["particles/m"]) == 0.5
    assert np.loadtxt(d / "00000_PS.txt").tolist() == [[1.0, 0.5], [2.0, 0.25]]
    info = (d / "00000_parameters.info").read_text(encoding="utf-8")
    assert "CurrentTime = 0.75" in info and "_hidden" not in info
    assert c["CurrentOutputNumber"] == 1
    with pytest.raises(ValueError):
        PM.analyzePM(gp, gd, p, dict(c, OutputFormat="jld"), dev, force=True)


def test_hdf5_output_layout_when_h5py_is_available(tmp_path):
    h5py = pytest.importorskip("h5py")     # not in the build image: the writer runs where the reference's HDF5 tooling exists
    gp = {1: PM.GridPatch([0, 0, 0], [1, 1, 1], 1, 1, (4, 4, 1))}
    gd = {1: PM.GridData({"ρD": np.full((1, 4, 4), 2.0), "Φ": np.zeros((1, 4, 4))})}
    p = {"x": np.arange(6.0).reshape(3, 2), "v": -np.arange(6.0).reshape(3, 2), "m": 0.5}
    PM._write_hdf5(str(tmp_path / "o.h5"), gp, gd, p)
    with h5py.File(tmp_path / "o.h5") as f:
        assert sorted(f["grid1"]) == ["Φ", "ρD"] and sorted(f["particles"]) == ["m", "vx", "vy", "x", "y"]
        assert f["grid1/ρD"].attrs["vsz_range"].tolist() == [0, 0, 1, 1]


def test_profiling_on_collects_the_step_timers(monkeypatch, tmp_path):
    """ProfilingOn = true (src/NoName.jl:86-93): the per-phase CUDA-event times of every step are accumulated and written
    to ProfilingResultsFile by noname()"""
    import json

    class St(_FakeStats):
        def as_dict(self):
            return {"kernel_launches": 7, "ms": {"deposit": 1.0, "solve": 2.0, "push": 3.0}}

    class Dev(_FakeDev):
        def step_static(self, dt):
            self.calls.append((dt,))
            return St(self.maxv)
    fake = Dev(2.0)
    monkeypatch.setattr(PM, "_device_for", lambda gp, p, c: fake)
    monkeypatch.setattr(PM, "_sync_host", lambda *a, **k: None)
    monkeypatch.setitem(NoName.INITIALIZERS, "InitializeParticleSimulation",
                        lambda gp, gd, p, c: gp.update({1: PM.GridPatch([0, 0, 0], [1, 1, 1], 1, 1, (16, 16, 1))}))
    prof = str(tmp_path / "prof.json")
    c = dict(cf.DEFAULTS, Initializer="InitializeParticleSimulation", Evolve="evolvePM", StartTime=0.0, StopTime=1.0, StopCycle=4,
             InitialDt=1e-3, CourantFactor=0.5, MaximumDt=0.03, OutputDisabled=True, ParticleDimensions=[8, 8, 1],
             ProfilingOn=True, ProfilingResultsFile=prof)
    NoName.noname(c)
    out = json.load(open(prof))
    assert out["steps"] == 4 and out["kernel_launches"] == 28
    assert out["ms_total"] == {"deposit": 4.0, "solve": 8.0, "push": 12.0} and out["ms_per_step"]["push"] == 3.0
    c2 = dict(c, ProfilingOn=False, ProfilingResultsFile=str(tmp_path / "none.json"))
    c2.pop("_profile", None)
    NoName.noname(c2)
    assert not (tmp_path / "none.json").exists()


def test_restart_roundtrip_resumes_time(tmp_path):
    """src/restart.jl + IOtools.jl:118-138 (with the F10 arity bug fixed): state and clock survive."""
    gp = {1: PM.GridPatch([0, 0, 0], [1, 1, 1], 1, 1, (8, 8, 1))}
    gd = {1: PM.GridData({"ρD": np.ones((1, 8, 8)), "Φ": np.zeros((1, 8, 8))})}
    p = {"x": np.random.default_rng(0).random((10, 2)), "v": np.zeros((10, 2)), "m": 0.5}
    c = dict(cf.DEFAULTS, CurrentTime=1.25, CurrentCycle=7, TopgridDimensions=[8, 8, 1])
    f = str(tmp_path / "restart.bin")
    NoName.write_all(f, c, gp, gd, p)
    c2 = dict(cf.DEFAULTS, _restart_file=f, _rconf="")
    gp2, gd2, p2 = {}, {}, {}
    NoName.restart(gp2, gd2, p2, c2)
    assert c2["StartTime"] == 1.25 and c2["StartCycle"] == 7 and c2["_rank"] == 2
    assert np.array_equal(p2["x"], p["x"]) and p2["m"] == 0.5 and gp2[1].dim == (8, 8, 1)
