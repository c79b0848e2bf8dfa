"""The reference's particle-mesh operator / plugin surface (src/ParticleMesh.jl), backed by the
CUDA library.  Same names, argument order and in-place semantics as the Julia functions:

    make_periodic(x)                                           :409
    deposit(rho, x, m, interpolation="none"|"cic")             :289
    compute_potential(phi, rho, rt)                            :417
    compute_acceleration(a, phi)                               :631
    interpVecToPoints(pa, a, x, interpolation="none"|"cic")    :640
    drift(x, v, dt)                                            :723
    kick(v, a, dt)  /  kick(v, pa, dt, a, dadt)                :731 / :800
    powerSpectrum(gp, gd, p)                                   :65
    evolvePM(gp, gd, p) / evolvePMCosmology(gp, gd, p)         :738 / :820   (conf["Evolve"] plugins)
    InitializeParticleSimulation + ProblemDefinition plugins   :103-287

The operator functions take HOST numpy arrays in the reference's memory layout (numpy C order with
the Julia axes reversed: rho[k,j,i], acc[k,j,i,c], x[n,d]), copy them to the GPU, run the device
operator and copy the result back into the first argument -- they exist for drop-in use and for
field-level parity tests.  The evolve loops keep all state on the device between steps and only
copy back at outputs (analyzePM) and at the end.  No CPU arithmetic fallback exists: without the
built library or without a GPU every call raises.
"""
from __future__ import annotations

import math
import os

import numpy as np

from . import cosmology as _cos
from . import _lib
from .GridTypes import GridPatch, GridData
from .device import DevicePM

_INTERP = ("none", "cic")
_cache = {}
#: module-level options applied to contexts created by the operator functions / evolve loops
options = {"device": 0, "deposit_mode": "atomic", "bugcompat_interp_z": False, "sort_every": 0}


def _check_interp(interpolation):
    if interpolation == "sic":
        raise NotImplementedError("interpolation='sic' is dead code in the reference (ParticleMesh.jl:298, "
                                  "sicdensity indexes undefined names); not provided")
    if interpolation not in _INTERP:
        raise ValueError(f"interpolation must be 'none' or 'cic', got {interpolation!r}")


def _ctx(dims, npart, rank, dep="cic", back="cic"):
    key = (tuple(dims), int(npart), int(rank), dep, back, options["device"], options["deposit_mode"],
           bool(options["bugcompat_interp_z"]))
    c = _cache.get(key)
    if c is None:
        if len(_cache) >= 4:
            _cache.pop(next(iter(_cache))).close()
        c = DevicePM(dims, npart, rank=rank, deposit=dep, backinterp=back, device=options["device"],
                     deposit_mode=options["deposit_mode"], bugcompat_interp_z=options["bugcompat_interp_z"])
        _cache[key] = c
    return c


def clear_cache():
    while _cache:
        _cache.popitem()[1].close()


def _mesh_dims(a):
    if a.ndim not in (3, 4):
        raise ValueError("mesh arrays must be (N3,N2,N1) or (N3,N2,N1,3)")
    n3, n2, n1 = a.shape[:3]
    return (n1, n2, n3)


def _rank_of(dims):
    return sum(1 for d in dims if d > 1)


# ---------------------------------------------------------------------------------------------
# operators
# ---------------------------------------------------------------------------------------------
def make_periodic(x):
    """:409-415"""
    r = x.shape[1]
    c = _ctx((4, 4, 4 if r == 3 else 1), x.shape[0], r)
    c.upload(x, None, 1.0)
    c.make_periodic()
    c.download(x, np.empty_like(x))
    return None


def deposit(rho, x, m, interpolation="none"):
    """:289-303 (zero fill + ngpdensity / cicdensity)"""
    _check_interp(interpolation)
    c = _ctx(_mesh_dims(rho), x.shape[0], x.shape[1], dep=interpolation)
    c.upload(x, None, m)
    c.deposit()
    c.get_field("rho", rho)
    return None


def compute_potential(phi, rho, rt=None):
    """:417-421 -> potential_2d/3d.  `rt` is the reference's caller-allocated rfft scratch; the
    device keeps its own spectrum buffer, so rt is accepted and left untouched."""
    dims = _mesh_dims(rho)
    if phi.shape != rho.shape:
        raise ValueError("phi and rho must have the same shape")
    c = _ctx(dims, 1, _rank_of(dims))
    c.set_field("rho", rho)
    c.solve()
    c.get_field("phi", phi)
    return None


def compute_acceleration(a, phi):
    """:631-637 (deriv1 with fac = Ng1 for every axis, then negate)"""
    dims = _mesh_dims(phi)
    c = _ctx(dims, 1, _rank_of(dims))
    c.set_field("phi", phi)
    c.accel()
    c.get_field("acc", a)
    return None


def interpVecToPoints(pa, a, x, interpolation="none"):
    """:640-663"""
    _check_interp(interpolation)
    c = _ctx(_mesh_dims(a), x.shape[0], x.shape[1], back=interpolation)
    c.upload(x, None, 1.0)
    c.set_field("acc", a)
    c.interp()
    c.get_field("pa", pa)
    return None


def drift(x, v, dt):
    """:723-728"""
    r = x.shape[1]
    c = _ctx((4, 4, 4 if r == 3 else 1), x.shape[0], r)
    c.upload(x, v, 1.0)
    c.drift(dt)
    c.download(x, np.empty_like(v))
    return None


def kick(v, a, dt, ascale=None, dadt=None):
    """static :731-736  /  comoving semi-implicit :800-807"""
    r = v.shape[1]
    c = _ctx((4, 4, 4 if r == 3 else 1), v.shape[0], r)
    c.upload(v, v, 1.0)
    c.set_field("pa", a)
    c.kick(dt, ascale, dadt)
    c.download(np.empty_like(v), v)
    return None


def powerSpectrum(gp, gd, p, conf=None):
    """:65-100.  Returns karray, power."""
    x = p["x"]
    dims = tuple(gp[1].dim)
    interp = (conf or {}).get("ParticleDepositInterpolation", "cic")
    c = _ctx(dims, x.shape[0], x.shape[1], dep=interp)
    make_periodic(x)                       # the reference mutates p["x"] here (:67)
    c.upload(x, None, p["m"])
    return c.power_spectrum()


# ---------------------------------------------------------------------------------------------
# initial conditions                                                           :161-198, 257-287
# ---------------------------------------------------------------------------------------------
def initialize_particles_uniform(x, ParticleDimensions):
    """:257-287 -- lattice (i-1/2)/Npd, i fastest."""
    rank = x.shape[1]
    pd = [int(q) for q in ParticleDimensions]
    if rank == 2:
        jj, ii = np.meshgrid(np.arange(1, pd[1] + 1), np.arange(1, pd[0] + 1), indexing="ij")
        x[:, 0], x[:, 1] = ii.ravel(), jj.ravel()
    elif rank == 3:
        kk, jj, ii = np.meshgrid(np.arange(1, pd[2] + 1), np.arange(1, pd[1] + 1), np.arange(1, pd[0] + 1),
                                 indexing="ij")
        x[:, 0], x[:, 1], x[:, 2] = ii.ravel(), jj.ravel(), kk.ravel()
    else:
        raise ValueError("rank must be 2 or 3")
    x[...] = (x - 0.5) / np.asarray(pd[:rank], dtype=np.float64)
    return None


def UniformParticles(gp, gd, p, conf):
    initialize_particles_uniform(p["x"], conf["ParticleDimensions"])


def UniformParticleWind(gp, gd, p, conf):
    initialize_particles_uniform(p["x"], conf["ParticleDimensions"])
    p["v"][:, 0] = conf["UniformParticleWindVelocity"]


def ZeldovichPancakeParticles(gp, gd, p, conf):
    """:177-198"""
    zc = conf["ZeldovichCausticRedshift"]
    kZ = 1.0
    zinit = conf["InitialRedshift"]
    cosmo = conf["_cosmo"]
    A = 5.0 * (1.0 + zc) / (2 * kZ) * 1.0 / (1.0 + zinit)
    x, v = p["x"], p["v"]
    initialize_particles_uniform(x, conf["ParticleDimensions"])
    adot = _cos.dadt_from_t_internal(cosmo, conf["StartTime"], zinit)
    a = _cos.a_from_t_internal(cosmo, conf["StartTime"], zinit, zwherea1=zinit)
    v[:, 0] = 2.0 / 5 * A * adot * np.sin(2 * math.pi * kZ * x[:, 0])
    x[:, 0] += 2.0 / 5 * A * a * np.sin(2 * math.pi * kZ * x[:, 0])
    np.add(x, 1.0, out=x, where=x < 0.0)     # make_periodic, host side (initialisation only)
    np.subtract(x, 1.0, out=x, where=x > 1.0)


def SyntheticZeldovichParticles(gp, gd, p, conf):
    """Seeded plane-wave Zel'dovich ICs generated on the device (new: the reference's
    PowerSpectrumParticles is unseeded and does not parse).  conf keys:
    SyntheticSeed, SyntheticModes, SyntheticKmax, SyntheticDispRms (box units), SyntheticVfac."""
    p["_synthetic"] = dict(seed=int(conf.get("SyntheticSeed", 12345)), nmodes=int(conf.get("SyntheticModes", 32)),
                           kmax=int(conf.get("SyntheticKmax", 8)),
                           disp_rms=float(conf.get("SyntheticDispRms", 0.3 / max(conf["TopgridDimensions"]))),
                           vfac=float(conf.get("SyntheticVfac", 0.0)))


PROBLEMS = {f.__name__: f for f in (UniformParticles, UniformParticleWind, ZeldovichPancakeParticles,
                                    SyntheticZeldovichParticles)}


def InitializeParticleSimulation(gp, gd, p, conf):
    """:103-159"""
    tgd = [int(d) for d in conf["TopgridDimensions"]]
    rank = sum(1 for d in tgd if d > 1)
    conf["_rank"] = rank
    npart = int(np.prod([int(q) for q in conf["ParticleDimensions"]]))
    dim = tuple(tgd)
    gp[1] = GridPatch(conf["DomainLeftEdge"], conf["DomainRightEdge"], 1, 1, dim)
    gd[1] = GridData({})
    shape = (dim[2], dim[1], dim[0])
    gd[1].d["ρD"] = np.ones(shape)
    gd[1].d["Φ"] = np.ones(shape)
    lazy = conf.get("ProblemDefinition") == "SyntheticZeldovichParticles"
    p["x"] = None if lazy else np.zeros((npart, rank))
    p["v"] = None if lazy else np.zeros((npart, rank))
    p["m"] = 1.0 * float(np.prod(dim)) / npart
    if "Cosmology" in conf:
        cosmo = _cos.parse_cosmology(conf["Cosmology"])
        conf["_cosmo"] = cosmo
        if "InitialRedshift" in conf:
            zstart = conf["InitialRedshift"]
            u = _cos.get_units(cosmo, zstart, zinit=zstart)
            tstart = cosmo.age_gyr(zstart) * (_cos.gyr_in_s / u.T)
            tend = cosmo.age_gyr(conf["FinalRedshift"]) * (_cos.gyr_in_s / u.T)
            conf["CurrentTime"] = tstart
            conf["CurrentRedshift"] = zstart
            conf["StartTime"] = tstart
            conf["StopTime"] = tend
            conf["OmegaCDM"] = cosmo.Om - conf.get("OmegaBaryon", 0.0)
            p["m"] *= conf["OmegaCDM"]
    name = conf["ProblemDefinition"]
    if name == "PowerSpectrumParticles":
        raise NotImplementedError(
            "PowerSpectrumParticles does not parse in the reference (ParticleMesh.jl:246) and draws unseeded "
            "randn; use ProblemDefinition = SyntheticZeldovichParticles (seeded, device-generated)")
    if name not in PROBLEMS:
        raise KeyError(f"unknown ProblemDefinition {name!r}")
    PROBLEMS[name](gp, gd, p, conf)
    return None


# ---------------------------------------------------------------------------------------------
# output hook                                                                      :13-63
# ---------------------------------------------------------------------------------------------
def readyForOutput(c):
    """src/IOtools.jl:40-59"""
    output = False
    if c["OutputEveryNCycle"] > 0 and c["CurrentCycle"] % c["OutputEveryNCycle"] == 0:
        output = True
    if c["OutputDtDataDump"] != 0 and abs(c["CurrentTime"] - c["LastOutputTime"]) >= c["OutputDtDataDump"]:
        output = True
    if output:
        c["LastOutputTime"] = c["CurrentTime"]
        c["LastOutputCycle"] = c["CurrentCycle"]
    return output


def _sync_host(dev, gd, p, fields=True):
    """Device -> host copy of the state the reference keeps in gd / p (only at outputs)."""
    n = dev.local_count
    if p.get("x") is None or p["x"].shape[0] != n:
        p["x"] = np.empty((n, dev.rank))
        p["v"] = np.empty((n, dev.rank))
    dev.download(p["x"], p["v"])
    if fields:
        try:
            dev.get_field("phi", gd[1].d["Φ"])
        except _lib.NNPMError:
            pass


def analyzePM(gp, gd, p, conf, dev=None, force=False):
    """:13-63.  HDF5/Winston are not available to this host; datasets with the reference's names
    (grid1/ρD, grid1/Φ, particles/x, vx, ...) go to NNNNN.npz beside NNNNN_parameters.info."""
    if not readyForOutput(conf) and not force:
        return None
    if conf.get("OutputDisabled"):
        return None
    c = conf
    if dev is not None:
        _sync_host(dev, gd, p)
    os.makedirs(c["OutputDirectory"], exist_ok=True)
    on = c["CurrentOutputNumber"]
    s = "%05d" % on
    newd = c["OutputDirectory"]
    if c["OutputSeparateDirectories"]:
        newd = os.path.join(newd, c.get("OutputDirPrefix", "") + s)
        os.makedirs(newd, exist_ok=True)
    with open(os.path.join(newd, s + "_parameters.info"), "w", encoding="utf-8") as f:
        for k, v in c.items():
            if not k.startswith("_"):
                f.write(f"{k} = {v}\n")
    c["CurrentOutputNumber"] += 1
    out = {"grid1/ρD": np.squeeze(gd[1].d["ρD"]), "grid1/Φ": np.squeeze(gd[1].d["Φ"]), "particles/m": p["m"]}
    for d, nm in enumerate("xyz"[: p["x"].shape[1]]):
        out[f"particles/{nm}"] = p["x"][:, d]
        out[f"particles/v{nm}"] = p["v"][:, d]
    np.savez(os.path.join(newd, s + ".npz"), **out)
    return None


# ---------------------------------------------------------------------------------------------
# evolve loops (conf["Evolve"] plugins)
# ---------------------------------------------------------------------------------------------
def _device_for(gp, p, conf):
    dims = tuple(gp[1].dim)
    rank = conf.get("_rank", sum(1 for d in dims if d > 1))
    npart = int(np.prod([int(q) for q in conf["ParticleDimensions"]]))
    dev = DevicePM(dims, npart, rank=rank, deposit=conf["ParticleDepositInterpolation"],
                   backinterp=conf["ParticleBackInterpolation"],
                   deposit_mode=conf.get("DepositMode", options["deposit_mode"]),
                   sort_every=int(conf.get("SortEvery", options["sort_every"])),
                   bugcompat_interp_z=bool(conf.get("BackInterpBugCompat", options["bugcompat_interp_z"])),
                   device=int(conf.get("Device", options["device"])))
    if p.get("_synthetic") is not None:
        dev.init_synthetic(conf["ParticleDimensions"], m=p["m"], **p["_synthetic"])
    else:
        dev.upload(p["x"], p["v"], p["m"])
    return dev


def evolvePM(gp, gd, p, conf, log=None):
    """:738-796 -- static PM loop; dt control on the host from the step's device reductions."""
    c = conf
    dev = _device_for(gp, p, c)
    try:
        dx = 1.0 / max(gp[1].dim)
        dt = c["InitialDt"]
        c["CurrentTime"] = c["StartTime"]
        c["CurrentCycle"] = c["StartCycle"]
        while c["CurrentTime"] < c["StopTime"] and c["CurrentCycle"] < c["StopCycle"]:
            st = dev.step_static(dt)
            c["CurrentTime"] += dt
            if log is not None:
                log.append((c["CurrentTime"], dt, st.max_pa, st.max_v))
            dt = c["CourantFactor"] * dx / (st.max_abs_v + 1e-30)
            dt = min(dt, c["MaximumDt"])
            if c["CurrentTime"] + dt > c["StopTime"]:
                dt = (1.0 + 1e-15) * (c["StopTime"] - c["CurrentTime"])
                if c.get("verbose"):
                    print("final dt = ", dt)
            c["CurrentCycle"] += 1
            analyzePM(gp, gd, p, c, dev)
        _sync_host(dev, gd, p)
    finally:
        dev.close()
    return None


def evolvePMCosmology(gp, gd, p, conf, log=None):
    """:820-894 -- comoving loop."""
    c = conf
    cosmo = c["_cosmo"]
    dev = _device_for(gp, p, c)
    try:
        dx = 1.0 / max(gp[1].dim)
        dt = c["InitialDt"]
        zinit = c["InitialRedshift"]
        c["CurrentTime"] = c["StartTime"]
        c["CurrentCycle"] = c["StartCycle"]
        c["LastOutputTime"] = c["CurrentTime"] - c["OutputDtDataDump"]
        while c["CurrentTime"] < c["StopTime"] and c["CurrentCycle"] < c["StopCycle"]:
            t = c["CurrentTime"]
            a1 = _cos.a_from_t_internal(cosmo, t + dt / 2 / 2, zinit, zwherea1=zinit)
            ak = _cos.a_from_t_internal(cosmo, t + dt / 2, zinit, zwherea1=zinit)
            adk = _cos.dadt(cosmo, ak, zwherea1=zinit)
            a2 = _cos.a_from_t_internal(cosmo, t + 3.0 * dt / 2 / 2, zinit, zwherea1=zinit)
            st = dev.step_comoving(dt, a1, ak, adk, a2)
            analyzePM(gp, gd, p, c, dev)
            c["CurrentTime"] += dt
            c["CurrentRedshift"] = _cos.z_from_t_internal(cosmo, c["CurrentTime"], zinit)
            if c.get("verbose"):
                print("max pa:", st.max_pa, " max v:", st.max_v, " dt:", dt, " z=", c["CurrentRedshift"])
            if log is not None:
                log.append((c["CurrentTime"], dt, st.max_pa, st.max_v))
            dt = c["CourantFactor"] * dx / (st.max_abs_v + 1e-30)
            dt = min(dt, c["DtFractionOfTime"] * c["CurrentTime"], c["MaximumDt"])
            if c["CurrentTime"] + dt > c["StopTime"]:
                dt = (1.0 + 1e-6) * (c["StopTime"] - c["CurrentTime"])
                if c.get("verbose"):
                    print("final dt = ", dt)
            c["CurrentCycle"] += 1
            analyzePM(gp, gd, p, c, dev)
        analyzePM(gp, gd, p, c, dev, force=True)
        _sync_host(dev, gd, p)
    finally:
        dev.close()
    return None


EVOLVERS = {"evolvePM": evolvePM, "evolvePMCosmology": evolvePMCosmology}
