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
import re

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


def _pancake_amplitudes(conf):
    """:178-191: A = 5 (1 + zc) / (2 kZ) / (1 + zinit); x += 2/5 A a sin, v = 2/5 A adot sin."""
    zc, kZ, zinit = conf["ZeldovichCausticRedshift"], 1.0, conf["InitialRedshift"]
    cosmo = conf["_cosmo"]
    A = 5.0 * (1.0 + zc) / (2 * kZ) * 1.0 / (1.0 + zinit)
    adot = _cos.dadt_from_t_internal(cosmo, conf["StartTime"], zinit)
    a = _cos.a_from_t_internal(cosmo, conf["StartTime"], zinit, zwherea1=zinit)
    return 2.0 / 5 * A * a, 2.0 / 5 * A * adot


def ZeldovichPancakeParticles(gp, gd, p, conf):
    """:177-198 (host arrays, as the reference; DeviceInitialConditions = true generates the same on the device)"""
    xamp, vamp = _pancake_amplitudes(conf)
    x, v = p["x"], p["v"]
    initialize_particles_uniform(x, conf["ParticleDimensions"])
    sn = np.sin(2 * math.pi * 1.0 * x[:, 0])
    v[:, 0] = vamp * sn
    x[:, 0] += xamp * sn
    np.add(x, 1.0, out=x, where=x < 0.0)     # make_periodic, host side (initialisation only)
    np.subtract(x, 1.0, out=x, where=x > 1.0)


def SyntheticZeldovichParticles(gp, gd, p, conf):
    """Seeded plane-wave Zel'dovich ICs generated on the device (test workload; not in the reference).  conf keys:
    SyntheticSeed, SyntheticModes, SyntheticKmax, SyntheticDispRms (box units), SyntheticVfac."""
    p["_ic"] = ("synthetic", dict(seed=int(conf.get("SyntheticSeed", 12345)), nmodes=int(conf.get("SyntheticModes", 32)),
                                  kmax=int(conf.get("SyntheticKmax", 8)),
                                  disp_rms=float(conf.get("SyntheticDispRms", 0.3 / max(conf["TopgridDimensions"]))),
                                  vfac=float(conf.get("SyntheticVfac", 0.0))))


def _parse_input_power_spectrum(expr):
    """InputPowerSpectrum = scaleFreePS(n, norm) | multiplePeaksPS(n, norm, [k1, k2, ...], lnwidth | [w1, w2, ...])
    (src/InputPowerSpectra.jl:1-24; doc/InputPowerSpectra.md) -> (n, norm, peaks) with peaks = None or (k_list, w_list);
    a scalar width applies to every peak (:20-24)."""
    from .conf import parse_value, _split_top
    m = re.match(r"^\s*(scaleFreePS|multiplePeaksPS)\s*\((.*)\)\s*$", str(expr))
    if not m:
        raise NotImplementedError(f"InputPowerSpectrum = {expr!r}: scaleFreePS(n, norm) and multiplePeaksPS(n, norm, k, lnwidth) are provided")
    args = [parse_value(a) for a in _split_top(m.group(2))]
    num = (int, float)
    if m.group(1) == "scaleFreePS":
        if len(args) != 2 or not all(isinstance(a, num) for a in args):
            raise ValueError(f"scaleFreePS(n, norm) takes two numbers: {expr!r}")
        return float(args[0]), float(args[1]), None
    if len(args) != 4 or not all(isinstance(a, num) for a in args[:2]) or not isinstance(args[2], list):
        raise ValueError(f"multiplePeaksPS(n, norm, [k...], lnwidth) expected: {expr!r}")
    ks = [float(q) for q in args[2]]
    ws = [float(q) for q in args[3]] if isinstance(args[3], list) else [float(args[3])] * len(ks)
    if len(ws) != len(ks) or not ks:
        raise ValueError(f"multiplePeaksPS: one width per peak (or a single number): {expr!r}")
    return float(args[0]), float(args[1]), (ks, ws)


def PowerSpectrumParticles(gp, gd, p, conf):
    """:209-254 with InputPowerSpectrum = scaleFreePS(n, norm) or multiplePeaksPS(n, norm, k, lnwidth), generated on the
    device by nnpm_init_grf / nnpm_init_grf_peaks -- seeded (conf RandomSeed, default 12345), rescaled to a per-component
    RMS displacement of DisplacementRMS (box units, default 0.3 cell) and with growing-mode velocities v = (adot/a) psi
    when a cosmology is set: the reference's own function does not parse (:246), draws unseeded randn and leaves the
    velocities unimplemented (:251)."""
    if "InputPowerSpectrum" not in conf:
        raise KeyError("Requested to use PowerSpectrumParticles but did not specify InputPowerSpectrum")   # :212-216
    n, _norm, peaks = _parse_input_power_spectrum(conf["InputPowerSpectrum"])
    vfac = 0.0
    if "_cosmo" in conf and "InitialRedshift" in conf:
        zi = conf["InitialRedshift"]
        a = _cos.a_from_t_internal(conf["_cosmo"], conf["StartTime"], zi, zwherea1=zi)
        vfac = _cos.dadt_from_t_internal(conf["_cosmo"], conf["StartTime"], zi) / a
    kw = dict(seed=int(conf.get("RandomSeed", 12345)), ps_index=n,
              disp_rms=float(conf.get("DisplacementRMS", 0.3 / max(conf["TopgridDimensions"]))), vfac=vfac)
    if peaks is not None:
        kw["peaks"] = peaks
    p["_ic"] = ("grf", kw)


PROBLEMS = {f.__name__: f for f in (UniformParticles, UniformParticleWind, ZeldovichPancakeParticles,
                                    SyntheticZeldovichParticles, PowerSpectrumParticles)}
#: ProblemDefinitions whose particles can be generated on the device (conf DeviceInitialConditions = true): nothing
#: is allocated on the host until the first output -- at 1024^3 the host arrays alone are 51.5 GB
_DEVICE_IC = {"UniformParticles", "ZeldovichPancakeParticles", "SyntheticZeldovichParticles", "PowerSpectrumParticles"}


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
    name = conf.get("ProblemDefinition")
    always_device = name in ("SyntheticZeldovichParticles", "PowerSpectrumParticles")
    lazy = always_device or (bool(conf.get("DeviceInitialConditions", False)) and name in _DEVICE_IC)
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
    if name not in PROBLEMS:
        raise KeyError(f"unknown ProblemDefinition {name!r}")
    if lazy and not always_device:
        if name == "UniformParticles":
            p["_ic"] = ("lattice", {})
        else:
            xamp, vamp = _pancake_amplitudes(conf)
            p["_ic"] = ("pancake", dict(xamp=xamp, vamp=vamp))
    else:
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
    """Device -> host copy of the state the reference keeps in gd / p (only at outputs and at the end).
    gd[1].d["ρD"] receives the density of the last step's deposit, as in the reference, where `deposit` writes gd's array
    in place (:767, :857) and the solve works on a temporary: the evolve loops set the `keep_rho` option for the steps
    after which an output is due.  If no kept density exists (state uploaded but never stepped) the particles are
    deposited where they are.  With NGPUs > 1 every rank receives the complete arrays."""
    n = dev.local_count
    multi = dev.nranks > 1
    if multi:
        from . import slab
    if not multi:
        if p.get("x") is None or p["x"].shape[0] != n:
            p["x"] = np.empty((n, dev.rank))
            p["v"] = np.empty((n, dev.rank))
        dev.download(p["x"], p["v"])
    else:
        xl, vl, ids = dev.download(ids=True)
        ids = slab.allgather_rows(ids)
        xs, vs = slab.allgather_rows(xl), slab.allgather_rows(vl)
        if p.get("x") is None or p["x"].shape != xs.shape:
            p["x"], p["v"] = np.empty_like(xs), np.empty_like(vs)
        p["x"][ids] = xs            # the reference's particle order
        p["v"][ids] = vs
    if fields:
        def fetch(which, key):
            try:
                a = dev.get_field(which)
            except _lib.NNPMError:
                return False
            gd[1].d[key][...] = slab.allgather_slabs(a) if multi else a
            return True
        fetch("phi", "Φ")
        if not fetch("rho", "ρD"):
            dev.deposit()           # overwrites the mesh: phi was fetched first
            fetch("rho", "ρD")


def _output_due(c, t, cycle):
    """readyForOutput (src/IOtools.jl:40-59) evaluated WITHOUT its side effects at a future (time, cycle)."""
    if c.get("OutputDisabled"):
        return False
    if c["OutputEveryNCycle"] > 0 and cycle % c["OutputEveryNCycle"] == 0:
        return True
    return c["OutputDtDataDump"] != 0 and abs(t - c["LastOutputTime"]) >= c["OutputDtDataDump"]


def _write_hdf5(fname, gp, gd, p):
    """IOtools.grid_output + particle_output (src/IOtools.jl:73-116): group grid<id> with every gd array squeezed (and
    the vsz_range attribute veusz reads), group particles with x, vx[, y, vy[, z, vz]], m."""
    import h5py
    with h5py.File(fname, "w") as f:
        for k, cgp in gp.items():
            g = f.create_group(f"grid{cgp.id}")
            for name, a in gd[cgp.id].d.items():
                # numpy C order with the Julia axes reversed == the bytes HDF5.jl writes for the column-major array
                ds = g.create_dataset(name, data=np.squeeze(a))
                ds.attrs["vsz_range"] = [gp[1].LE[0], gp[1].LE[1], gp[1].RE[0], gp[1].RE[1]]
        g = f.create_group("particles")
        for d, nm in enumerate("xyz"[: p["x"].shape[1]]):
            g[nm] = np.ascontiguousarray(p["x"][:, d])
            g[f"v{nm}"] = np.ascontiguousarray(p["v"][:, d])
        g["m"] = p["m"]


def analyzePM(gp, gd, p, conf, dev=None, force=False):
    """:13-63.  NNNNN_parameters.info beside the data file: NNNNN.h5 with the reference's group / dataset names
    (grid1/ρD, grid1/Φ, particles/x, vx, ..., m) when h5py is importable, else NNNNN.npz under the same names (conf
    OutputFormat = hdf5 | npz forces one).  OutputPowerSpectrum = true (:47-61 saves a Winston plot) writes the numbers
    instead: NNNNN_PS.txt, columns k and P(k) from the device-side powerSpectrum."""
    if not readyForOutput(conf) and not force:
        return None
    if conf.get("OutputDisabled"):
        return None
    c = conf
    if dev is not None:
        _sync_host(dev, gd, p)
    pk = None
    if c.get("OutputPowerSpectrum") is True:
        # after the sync: the device-side P(k) re-deposits into the mesh (the next step deposits again anyway)
        pk = dev.power_spectrum() if dev is not None else powerSpectrum(gp, gd, p, c)
    if dev is not None and dev.nranks > 1 and dev.rank_id != 0:
        c["CurrentOutputNumber"] += 1     # every rank keeps the same counters; rank 0 writes the files
        return None
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
    fmt = str(c.get("OutputFormat", "auto")).lower()
    if fmt not in ("auto", "hdf5", "npz"):
        raise ValueError(f"OutputFormat must be hdf5, npz or auto (got {fmt!r})")
    if fmt == "auto":
        try:
            import h5py  # noqa: F401
            fmt = "hdf5"
        except ImportError:
            fmt = "npz"
    if fmt == "hdf5":
        _write_hdf5(os.path.join(newd, s + ".h5"), gp, gd, p)
    else:
        out = {"grid1/ρD": np.squeeze(gd[1].d["ρD"]), "grid1/Φ": np.squeeze(gd[1].d["Φ"]), "particles/m": p["m"]}
        for d, nm in enumerate("xyz"[: p["x"].shape[1]]):
            out[f"particles/{nm}"] = p["x"][:, d]
            out[f"particles/v{nm}"] = p["v"][:, d]
        np.savez(os.path.join(newd, s + ".npz"), **out)
    if pk is not None:
        np.savetxt(os.path.join(newd, s + "_PS.txt"), np.column_stack(pk), header=f"k P(k)   time = {c['CurrentTime']}")
    return None


# ---------------------------------------------------------------------------------------------
# evolve loops (conf["Evolve"] plugins)
# ---------------------------------------------------------------------------------------------
def _device_for(gp, p, conf):
    """Device context(s) for an evolve loop.  New conf keys (no reference counterpart): NGPUs (slab decomposition
    over that many processes / GPUs; the process must have been launched once per GPU), SortEvery, DepositMode =
    atomic | fixedpoint, BackInterpBugCompat, Device, PartCapacityFactor.  ProfilingOn = true (src/NoName.jl:86-93 wraps the run
    in Julia's sampling profiler) turns on the per-phase CUDA-event timers of the step calls instead (_profile_step)."""
    dims = tuple(gp[1].dim)
    rank = conf.get("_rank", sum(1 for d in dims if d > 1))
    npart = int(np.prod([int(q) for q in conf["ParticleDimensions"]]))
    ngpus = int(conf.get("NGPUs", 1))
    rid, local = 0, int(conf.get("Device", options["device"]))
    if ngpus > 1:
        from . import slab
        rid, _, local = slab.require_world(ngpus)
    cap = 0
    if ngpus > 1 and "PartCapacityFactor" in conf:
        cap = int(float(conf["PartCapacityFactor"]) * npart / ngpus) + 4096
    dev = DevicePM(dims, npart, rank=rank, deposit=conf["ParticleDepositInterpolation"],
                   backinterp=conf["ParticleBackInterpolation"],
                   deposit_mode=conf.get("DepositMode", options["deposit_mode"]),
                   sort_every=int(conf.get("SortEvery", options["sort_every"])),
                   bugcompat_interp_z=bool(conf.get("BackInterpBugCompat", options["bugcompat_interp_z"])),
                   device=local, nranks=ngpus, rank_id=rid, part_capacity=cap, timing=bool(conf.get("ProfilingOn", False)))
    if ngpus > 1:
        dev.comm_init(slab.broadcast_bytes(DevicePM.comm_unique_id() if rid == 0 else None))
    if p.get("_ic") is not None and p.get("x") is None:
        kind, kw = p["_ic"]
        pd = conf["ParticleDimensions"]
        if kind == "grf":
            dev.init_grf(pd, m=p["m"], **kw)
        elif kind == "pancake":
            dev.init_pancake(pd, m=p["m"], **kw)
        elif kind == "lattice":
            dev.init_synthetic(pd, kind="lattice", m=p["m"])
        else:
            dev.init_synthetic(pd, m=p["m"], **kw)
    elif ngpus > 1:
        # every rank holds the full host arrays (the initialisers ran on each): upload one contiguous share, then
        # route the particles to their slabs
        lo, hi = rid * npart // ngpus, (rid + 1) * npart // ngpus
        dev.upload(np.ascontiguousarray(p["x"][lo:hi]), np.ascontiguousarray(p["v"][lo:hi]), p["m"], first_id=lo)
        dev.redistribute()
    else:
        dev.upload(p["x"], p["v"], p["m"])
    return dev


def _step_stats(dev, st):
    return dev.reduce_stats(st) if dev.nranks > 1 else st


def _profile_step(conf, st):
    """ProfilingOn: accumulate the step's per-phase CUDA-event milliseconds (nnpm_stats.ms) and launch count in
    conf["_profile"]; NoName.noname() writes them to ProfilingResultsFile (the reference serialises Profile.retrieve()
    there, src/IOtools.jl:142-147)."""
    if not conf.get("ProfilingOn", False):
        return
    prof = conf.setdefault("_profile", {"steps": 0, "kernel_launches": 0, "ms": {}})
    d = st.as_dict()
    prof["steps"] += 1
    prof["kernel_launches"] += int(d["kernel_launches"])
    for k, v in d["ms"].items():
        prof["ms"][k] = prof["ms"].get(k, 0.0) + float(v)


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
            # keep the deposited density only for steps after which analyzePM will write (or the loop ends)
            t1, cyc1 = c["CurrentTime"] + dt, c["CurrentCycle"] + 1
            dev.set_option("keep_rho", _output_due(c, t1, cyc1) or not (t1 < c["StopTime"] and cyc1 < c["StopCycle"]))
            st = _step_stats(dev, dev.step_static(dt))
            _profile_step(c, st)
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
            t1, cyc1 = t + dt, c["CurrentCycle"] + 1
            dev.set_option("keep_rho", _output_due(c, t, c["CurrentCycle"]) or _output_due(c, t1, cyc1)
                           or not (t1 < c["StopTime"] and cyc1 < c["StopCycle"]))
            st = _step_stats(dev, dev.step_comoving(dt, a1, ak, adk, a2))
            _profile_step(c, st)
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
