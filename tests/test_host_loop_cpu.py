"""The product's host loop end to end on CPU: run_noname() on the reference's shipped run/CosmologyPM conf (pancake ICs,
evolvePMCosmology's dt control, output cadence, keep_rho requests, restart file) with the device context replaced by an
adapter over the C restatement of the operators -- so that only HOST logic differs from the oracle's own loop, and the
two must agree in step count, final time (to the last bits) and state.  The same run on the GPU is
tests/test_gpu_run.py::test_run_noname_shipped_cosmology_conf."""
import os

import numpy as np
import pytest

from oracle import pm_ref as O
from test_gpu_run import SHIPPED_COSMOLOGY_PM, SHIPPED_PM


class _Stats:
    def __init__(self, a, b, c):
        self.max_abs_v, self.max_v, self.max_pa = a, b, c


def _adapter():
    from noname_b200 import _lib
    from oracle.pm_ref_c import PMRefC

    class OracleDev:
        """DevicePM's interface (the part the evolve loops and analyzePM use) over oracle/pm_ref.c, single process"""
        nranks, rank_id = 1, 0

        def __init__(self, dims, npart_total, rank=3, **kw):
            self.dims, self.rank = tuple(dims), rank
            self.f = O.Fields(self.dims, npart_total, rank)
            self.c = PMRefC(threads=1)
            self.keep, self.rho, self.steps_kept = False, None, 0
            self.dep, self.back = kw.get("deposit", "cic"), kw.get("backinterp", "cic")

        def set_option(self, name, value):
            assert name == "keep_rho"
            self.keep = bool(value)

        def upload(self, x, v, m, first_id=0):
            self.x, self.v, self.m = x.copy(), v.copy(), m

        @property
        def local_count(self):
            return len(self.x)

        def init_grf(self, pdims, seed=12345, ps_index=-1.0, disp_rms=0.0, vfac=0.0, m=1.0, peaks=None):
            self.x, self.v = O.synthetic_grf(tuple(int(q) for q in pdims), self.rank, seed, ps_index, disp_rms, vfac, peaks=peaks)
            self.m, OracleDev.ic_request = m, dict(seed=seed, ps_index=ps_index, disp_rms=disp_rms, vfac=vfac, peaks=peaks)

        def init_pancake(self, pdims, xamp, vamp, m=1.0):
            q = O.initialize_particles_uniform(tuple(int(k) for k in pdims), self.rank)
            sn = np.sin(2 * np.pi * q[:, 0])
            self.x, self.v, self.m = q.copy(), np.zeros_like(q), m
            self.v[:, 0] = vamp * sn
            self.x[:, 0] += xamp * sn
            O.make_periodic(self.x)
            OracleDev.ic_request = dict(kind="pancake", xamp=xamp, vamp=vamp)

        def init_synthetic(self, pdims, kind="zeldovich", m=1.0, **kw):
            assert kind == "lattice"
            self.x = O.initialize_particles_uniform(tuple(int(k) for k in pdims), self.rank)
            self.v, self.m = np.zeros_like(self.x), m
            OracleDev.ic_request = dict(kind="lattice")

        def step_static(self, dt):
            mabs, mv = self.c.step_static(self.f, self.x, self.v, self.m, dt, self.dep, self.back)
            self.rho = self.f.rho.copy() if self.keep else None
            self.steps_kept += self.keep
            return _Stats(mabs, mv, 0.0)

        def step_comoving(self, dt, a1, ak, adk, a2):
            mabs, mv = self.c.step_comoving(self.f, self.x, self.v, self.m, dt, a1, ak, adk, a2)
            self.rho = self.f.rho.copy() if self.keep else None
            self.steps_kept += self.keep
            return _Stats(mabs, mv, 0.0)

        def download(self, x=None, v=None, ids=False):
            x[...], v[...] = self.x, self.v
            return x, v

        def get_field(self, which):
            if which == "phi":
                return self.f.phi.copy()
            if self.rho is None:
                raise _lib.NNPMError(-6, "mesh does not currently hold rho")
            return self.rho.copy()

        def power_spectrum(self):
            return O.powerSpectrum(self.x.copy(), self.m, self.dims, self.dep)

        def deposit(self):
            raise AssertionError("an output was written after a step whose density the loop had not asked to keep")

        def close(self):
            OracleDev.last = self

    return OracleDev


@pytest.mark.parametrize("zfinal", [100.0, 40.0])
def test_run_noname_host_loop_equals_the_oracle_loop(monkeypatch, tmp_path, zfinal):
    from noname_b200 import ParticleMesh as PM, NoName, run_noname
    dev_cls = _adapter()
    monkeypatch.setattr(PM, "DevicePM", dev_cls)
    conf = tmp_path / "particle_mesh.conf"
    conf.write_text(SHIPPED_COSMOLOGY_PM, encoding="utf-8")
    over = {"OutputDirectory": str(tmp_path / "out"), "RestartFileName": str(tmp_path / "restart.bin"), "OutputFormat": "npz",
            "FinalRedshift": zfinal}
    gp, gd, p = run_noname(str(conf), overrides=over)
    c = NoName.conf

    oc = O.Cosmo(h=1.0, OmegaM=1.0, OmegaR=0.0)
    ts, te = O.start_stop_times(oc, 400, zfinal)
    a = O.a_from_t_internal(oc, ts, 400, zwherea1=400)
    ad = O.dadt_from_t_internal(oc, ts, 400)
    xr, vr = O.zeldovich_pancake((128, 128, 1), 2, 10.0, 400, a, ad)
    m = 64 * 64 / (128.0 * 128.0)
    rc = dict(StartTime=ts, StopTime=te, StartCycle=0, StopCycle=1000000, InitialDt=1e-4, CourantFactor=0.5, MaximumDt=1.0,
              DtFractionOfTime=0.1, InitialRedshift=400)
    t, cyc, f = O.evolvePMCosmology(xr, vr, m, (64, 64, 1), rc, oc)
    # the product's closed-form / Newton a(t) against the oracle's quadrature + root finder, through the whole dt control
    assert c["CurrentCycle"] == cyc and c["CurrentTime"] == pytest.approx(t, rel=1e-13)
    assert c["CurrentRedshift"] == pytest.approx(zfinal, abs=1e-4)            # final-step clamp (1 + 1e-6), :882-885
    assert np.max(np.abs(p["x"] - xr)) < 1e-12 and np.max(np.abs(p["v"] - vr)) < 1e-12 * np.max(np.abs(vr))
    assert np.max(np.abs(gd[1].d["Φ"] - f.phi)) <= 1e-12 * np.max(np.abs(f.phi))
    assert np.max(np.abs(gd[1].d["ρD"] - f.rho)) <= 1e-12 * np.max(np.abs(f.rho))
    # outputs: OutputDtDataDump = 3.5 with LastOutputTime = StartTime - 3.5 -> a dump right after the first step, then every
    # 3.5 time units, then the forced final one; each holds the density of the step it follows (the adapter raises otherwise)
    outs = sorted(os.path.join(dp, q) for dp, _, fs in os.walk(over["OutputDirectory"]) for q in fs if q.endswith(".npz"))
    assert len(outs) >= 2
    assert np.load(outs[0])["grid1/ρD"].sum() == pytest.approx(m * 128 * 128, rel=1e-12)
    assert np.array_equal(np.load(outs[-1])["grid1/ρD"], np.squeeze(gd[1].d["ρD"]))
    assert 0 < dev_cls.last.steps_kept < cyc                                    # kept only where needed
    assert os.path.exists(over["RestartFileName"])


def test_run_noname_shipped_pm_conf_verbatim_host_loop(monkeypatch, tmp_path):
    """run/PM/particle_mesh.conf as shipped (static evolvePM, InputPowerSpectrum = multiplePeaksPS(-1, 1, [0.1, 1, 10], 1.),
    100^2 particles on a 1000 x 1000 x 1 mesh): the conf parser, the IC request handed to the device and evolvePM's dt
    control / output cadence, on CPU.  GPU twin: tests/test_gpu_run.py::test_run_noname_shipped_pm_conf_verbatim."""
    from noname_b200 import ParticleMesh as PM, NoName, run_noname
    dev_cls = _adapter()
    monkeypatch.setattr(PM, "DevicePM", dev_cls)
    conf = tmp_path / "particle_mesh.conf"
    conf.write_text(SHIPPED_PM, encoding="utf-8")
    over = {"OutputDirectory": str(tmp_path / "out"), "RestartFileName": str(tmp_path / "restart.bin"), "OutputFormat": "npz"}
    gp, gd, p = run_noname(str(conf), overrides=over)
    c = NoName.conf
    assert dev_cls.ic_request == dict(seed=12345, ps_index=-1.0, disp_rms=0.3 / 1000, vfac=0.0, peaks=([0.1, 1.0, 10.0], [1.0, 1.0, 1.0]))
    xr, vr = O.synthetic_grf((100, 100, 1), 2, 12345, -1.0, 0.3 / 1000, 0.0, peaks=([0.1, 1.0, 10.0], [1.0, 1.0, 1.0]))
    m = 1000.0 * 1000.0 / 1e4
    rc = dict(StartTime=0.0, StopTime=1.0, StartCycle=0, StopCycle=1000000, InitialDt=1e-7, CourantFactor=0.2, MaximumDt=0.1)
    t, cyc, f = O.evolvePM(xr, vr, m, (1000, 1000, 1), rc)
    assert c["CurrentCycle"] == cyc and c["CurrentTime"] == pytest.approx(t, rel=1e-14) and p["m"] == m
    assert np.max(np.abs(p["x"] - xr)) < 1e-13 and np.max(np.abs(p["v"] - vr)) <= 1e-12 * np.max(np.abs(vr))
    assert np.max(np.abs(gd[1].d["ρD"] - f.rho)) <= 1e-12 * np.max(np.abs(f.rho))
    outs = [q for dp, _, fs in os.walk(over["OutputDirectory"]) for q in fs if q.endswith(".npz")]
    assert len(outs) == 1                      # OutputDtDataDump = 0.5: one dump at t = 0.5 + InitialDt, none at StopTime (:780-796)


def test_restart_continues_a_stopped_run(monkeypatch, tmp_path):
    """src/restart.jl through run_noname(restart=True): a run stopped by StopCycle leaves a restart file; restarting
    from it (with an rconf that lifts StopCycle) resumes at the dumped time and cycle, with the particles of the dump,
    and reaches StopTime"""
    from noname_b200 import ParticleMesh as PM, NoName, run_noname
    monkeypatch.setattr(PM, "DevicePM", _adapter())
    conf = tmp_path / "particle_mesh.conf"
    conf.write_text(SHIPPED_COSMOLOGY_PM, encoding="utf-8")
    rfile = str(tmp_path / "restart.bin")
    over = {"OutputDirectory": str(tmp_path / "out"), "RestartFileName": rfile, "OutputFormat": "npz", "FinalRedshift": 100.0,
            "StopCycle": 20}
    gp, gd, p = run_noname(str(conf), overrides=over)
    c1 = dict(NoName.conf)
    assert c1["CurrentCycle"] == 20 and c1["CurrentTime"] < c1["StopTime"]
    x20 = p["x"].copy()
    rconf = tmp_path / "restart.conf"
    rconf.write_text("StopCycle = 1000000\n", encoding="utf-8")
    gp2, gd2, p2 = run_noname(rfile, restart_=True, rconf=str(rconf), overrides={"OutputDirectory": str(tmp_path / "out2"),
                                                                                   "RestartFileName": str(tmp_path / "r2.bin")})
    c2 = NoName.conf
    assert c2["StartTime"] == c1["CurrentTime"] and c2["StartCycle"] == 20
    assert c2["CurrentCycle"] > 20 and c2["CurrentTime"] >= c2["StopTime"] == c1["StopTime"]
    assert c2["CurrentRedshift"] == pytest.approx(100.0, abs=1e-3)
    assert p2["x"].shape == x20.shape and not np.array_equal(p2["x"], x20)
    assert gd2[1].d["ρD"].sum() == pytest.approx(p2["m"] * 128 * 128, rel=1e-12)


def test_static_3d_wind_run_with_ngp_deposit_cycle_outputs_and_power_spectrum(monkeypatch, tmp_path):
    """a conf a NoName user could write: UniformParticleWind (:161-175), ParticleDepositInterpolation = none, CIC back-
    interpolation, OutputEveryNCycle, OutputPowerSpectrum -- the interpolation keys reach the device, every dump is
    preceded by a kept density, and the P(k) numbers are written beside it"""
    from noname_b200 import ParticleMesh as PM, NoName, run_noname
    dev_cls = _adapter()
    monkeypatch.setattr(PM, "DevicePM", dev_cls)
    conf = tmp_path / "wind.conf"
    conf.write_text("""
Initializer = InitializeParticleSimulation
ProblemDefinition = UniformParticleWind
UniformParticleWindVelocity = 0.25
Evolve = evolvePM
TopgridDimensions  = [16, 16, 16]
ParticleDimensions = [16, 16, 16]
MaximumDt = 0.05
StopTime = 0.4
CourantFactor = 0.5
ParticleDepositInterpolation = none
ParticleBackInterpolation    = cic
OutputEveryNCycle = 2
OutputPowerSpectrum = true
OutputSeparateDirectories = false
""", encoding="utf-8")
    over = {"OutputDirectory": str(tmp_path / "out"), "RestartFileName": str(tmp_path / "restart.bin"), "OutputFormat": "npz"}
    gp, gd, p = run_noname(str(conf), overrides=over)
    c = NoName.conf
    dev = dev_cls.last
    assert (dev.dep, dev.back) == ("none", "cic") and dev.dims == (16, 16, 16) and dev.rank == 3
    q = O.initialize_particles_uniform((16, 16, 16), 3)
    xr, vr = q.copy(), np.zeros_like(q)
    vr[:, 0] = 0.25
    rc = dict(StartTime=0.0, StopTime=0.4, StartCycle=0, StopCycle=1000000, InitialDt=1e-7, CourantFactor=0.5, MaximumDt=0.05)
    t, cyc, f = O.evolvePM(xr, vr, 1.0, (16, 16, 16), rc, dep="none", back="cic")      # evolves xr, vr in place
    assert c["CurrentCycle"] == cyc and c["CurrentTime"] == pytest.approx(t, rel=1e-14)
    assert np.max(np.abs(p["x"] - xr)) < 1e-14 and np.max(np.abs(p["v"] - vr)) < 1e-14
    # a uniform lattice in a uniform wind stays a lattice: no force, positions advance by 0.25 t (mod 1)
    assert np.max(np.abs(p["v"][:, 0] - 0.25)) < 1e-13 and np.max(np.abs(p["v"][:, 1:])) < 1e-13
    d = p["x"] - q
    d -= np.rint(d)
    assert np.max(np.abs(d[:, 0] - (0.25 * t - np.rint(0.25 * t)))) < 1e-12 and np.max(np.abs(d[:, 1:])) < 1e-13
    files = sorted(os.listdir(over["OutputDirectory"]))
    ndump = cyc // 2
    assert [q for q in files if q.endswith(".npz")] == ["%05d.npz" % i for i in range(ndump)]
    assert [q for q in files if q.endswith("_PS.txt")] == ["%05d_PS.txt" % i for i in range(ndump)]
    assert dev.steps_kept == ndump + (cyc % 2)            # one kept density per dump (+ the last step of the loop)


def test_north_star_aliases_and_device_initial_conditions(monkeypatch, tmp_path):
    """the north star's conf spellings -- dims, cosm_Ωm, cosm_h (absent from the reference's code, accepted
    as aliases) -- with DeviceInitialConditions = true: nothing is allocated on the host, the pancake request reaches the
    device with the reference's amplitudes (:178-191), and the host arrays appear at the first output"""
    from noname_b200 import ParticleMesh as PM, NoName, run_noname
    dev_cls = _adapter()
    monkeypatch.setattr(PM, "DevicePM", dev_cls)
    conf = tmp_path / "alias.conf"
    conf.write_text("""
Initializer = InitializeParticleSimulation
ProblemDefinition = ZeldovichPancakeParticles
ZeldovichCausticRedshift = 10.
Evolve = evolvePMCosmology
dims = [32, 32, 1]
ParticleDimensions = [64, 64, 1]
cosm_Ωm = 1.0
cosm_h = 1.0
InitialRedshift = 400
FinalRedshift = 200
InitialDt = 0.0001
CourantFactor = 0.5
DtFractionOfTime = 0.1
MaximumDt = 1.0
OmegaBaryon = 0.
DeviceInitialConditions = true
""", encoding="utf-8")
    over = {"OutputDirectory": str(tmp_path / "out"), "RestartFileName": str(tmp_path / "restart.bin"), "OutputFormat": "npz"}
    gp, gd, p = run_noname(str(conf), overrides=over)
    c = NoName.conf
    assert tuple(gp[1].dim) == (32, 32, 1) and c["Cosmology"] == "cosmology(h=1.0,OmegaM=1.0,OmegaR=0.0)"
    oc = O.Cosmo(h=1.0, OmegaM=1.0, OmegaR=0.0)
    ts, te = O.start_stop_times(oc, 400, 200)
    a = O.a_from_t_internal(oc, ts, 400, zwherea1=400)
    ad = O.dadt_from_t_internal(oc, ts, 400)
    A = 5.0 * 11.0 / 2.0 / 401.0
    assert dev_cls.ic_request["kind"] == "pancake"
    assert dev_cls.ic_request["xamp"] == pytest.approx(0.4 * A * a, rel=1e-12) and dev_cls.ic_request["vamp"] == pytest.approx(0.4 * A * ad, rel=1e-12)
    xr, vr = O.zeldovich_pancake((64, 64, 1), 2, 10.0, 400, a, ad)
    m = 32 * 32 / (64.0 * 64.0)
    rc = dict(StartTime=ts, StopTime=te, StartCycle=0, StopCycle=1000000, InitialDt=1e-4, CourantFactor=0.5, MaximumDt=1.0,
              DtFractionOfTime=0.1, InitialRedshift=400)
    t, cyc, f = O.evolvePMCosmology(xr, vr, m, (32, 32, 1), rc, oc)
    assert c["CurrentCycle"] == cyc and p["m"] == pytest.approx(m)
    assert p["x"].shape == (64 * 64, 2) and np.max(np.abs(p["x"] - xr)) < 1e-12 and np.max(np.abs(p["v"] - vr)) < 1e-12 * np.max(np.abs(vr))


def test_command_line_like_the_reference(monkeypatch, tmp_path):
    """IOtools.parse_commandline (src/IOtools.jl:21-38): positional configuration_file, --restart / -r, --rconf"""
    from noname_b200 import ParticleMesh as PM, NoName, run_noname
    monkeypatch.setattr(PM, "DevicePM", _adapter())
    conf = tmp_path / "cli.conf"
    conf.write_text(f"""
Initializer = InitializeParticleSimulation
ProblemDefinition = UniformParticles
Evolve = evolvePM
TopgridDimensions  = [8, 8, 1]
ParticleDimensions = [8, 8, 1]
MaximumDt = 0.1
StopTime = 0.3
OutputDirectory = "{tmp_path / 'out'}"
RestartFileName = "{tmp_path / 'restart.bin'}"
""", encoding="utf-8")
    gp, gd, p = run_noname(argv=[str(conf)])
    c1 = dict(NoName.conf)
    assert c1["CurrentTime"] >= 0.3 and c1["CurrentCycle"] >= 3 and os.path.exists(tmp_path / "restart.bin")
    rconf = tmp_path / "restart.conf"
    rconf.write_text("StopTime = 0.6\n", encoding="utf-8")
    gp, gd, p = run_noname(argv=["-r", "--rconf", str(rconf), str(tmp_path / "restart.bin")])
    c2 = NoName.conf
    assert c2["StartTime"] == c1["CurrentTime"] and c2["CurrentTime"] >= 0.6 and c2["CurrentCycle"] > c1["CurrentCycle"]
    assert np.allclose(gd[1].d["ρD"], 1.0)                      # a lattice at rest stays uniform


@pytest.mark.parametrize("text,path", [(SHIPPED_COSMOLOGY_PM, "run/CosmologyPM/particle_mesh.conf"), (SHIPPED_PM, "run/PM/particle_mesh.conf")])
def test_embedded_confs_are_the_reference_files_verbatim(text, path):
    """the tests carry the shipped confs as strings (the reference tree does not travel to the GPU box): where the tree is
    present they must be the files, byte for byte up to surrounding blank lines"""
    ref = os.path.join("/root/reference", path)
    if not os.path.exists(ref):
        pytest.skip("reference checkout not present")
    assert open(ref, encoding="utf-8").read().strip() == text.strip()


def test_builtin_defaults_equal_the_reference_default_conf():
    """noname_b200.conf.DEFAULTS restates src/default.conf (the first conf layer, src/NoName.jl:58): where the reference
    tree is present, parsing the file with the product's parser must give exactly those values"""
    from noname_b200 import conf as cf
    ref = "/root/reference/src/default.conf"
    if not os.path.exists(ref):
        pytest.skip("reference checkout not present")
    assert cf.parseconf(ref) == cf.DEFAULTS
