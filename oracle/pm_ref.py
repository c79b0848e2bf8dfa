"""CPU ORACLE (test infrastructure, not product code) -- numpy restatement of the
particle-mesh path of yipihey/NoName (`src/ParticleMesh.jl`, `src/cosmology.jl`).

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / --impl reference
legs may import this module.  Nothing under `noname_b200/` imports it.

PARITY UNPINNED: the reference ships no tests, golden vectors or fixtures,
cannot be run in this image (no Julia; Julia-0.4 syntax), and its FFT / cosmology
arithmetic lives in third-party packages that are not vendored (Julia Base FFTW,
Cosmology.jl, Roots.jl; all unpinned, Readme.md:38-49).  The oracle is therefore pinned by
(a) a literal pure-Python-loop transcription (`*_loops` functions below) that the vectorised
functions must match bit for bit, (b) the analytic known-answer tests in
`tests/test_oracle_kats.py`, and (c) the independent C restatement `oracle/pm_ref.c`.

Layout: Julia arrays are column-major.  Every array here is the SAME MEMORY viewed by
numpy in C order with the axes reversed:
    Julia rho[i,j,k]   (N1,N2,N3)      <->  numpy rho[k,j,i]   shape (N3,N2,N1)
    Julia acc[c,i,j,k] (3,N1,N2,N3)    <->  numpy acc[k,j,i,c] shape (N3,N2,N1,3)
    Julia x[d,n]       (rank,Npart)    <->  numpy x[n,d]       shape (Npart,rank)
    Julia rt[i,j,k]    (N1/2+1,N2,N3)  <->  numpy rt[k,j,i]    shape (N3,N2,N1//2+1)
2-D runs carry N3 == 1 and rank == 2 particle arrays (ParticleMesh.jl:105,122).
All indices below are 0-based (reference index minus one).
"""
from __future__ import annotations

import math

import numpy as np
import scipy.fft as sfft
from scipy import integrate, optimize

# --------------------------------------------------------------------------------------
# make_periodic                                                   ParticleMesh.jl:409-415
# --------------------------------------------------------------------------------------


def make_periodic(x):
    """x<0 -> x+1 ; x>1 -> x-1 (single wrap, x == 1.0 untouched). ParticleMesh.jl:409-415"""
    lo = x < 0.0
    x[lo] += 1.0
    hi = x > 1.0
    x[hi] -= 1.0
    return None


# --------------------------------------------------------------------------------------
# cell index / offset                                             ParticleMesh.jl:386-394
# --------------------------------------------------------------------------------------


def _cell(xd, n):
    """0-based cell index, +1 neighbour and offset of coordinate array xd on n cells.

    Reference (1-based): ii = floor(Int64, x*N)+1 ; ip1 = ii==N ? 1 : ii+1 ;
    dx = x*N - ii + 1   (evaluated left to right: (x*N - ii) + 1).
    Deviation, documented in SURVEY.md "x == 1.0 edge": the reference indexes N+1 under
    @inbounds when x == 1.0; here index N maps to 0 (the periodic intent, weight 0 on +1).
    """
    fn = float(n)
    t = xd * fn
    i1 = np.floor(t).astype(np.int64) + 1          # the reference's 1-based ii
    d = (t - i1.astype(np.float64)) + 1.0
    i0 = i1 - 1
    i0 = np.mod(i0, n)                              # x == 1.0 (and stray) -> periodic
    ip = np.where(i0 == n - 1, 0, i0 + 1)
    return i0, ip, d


# --------------------------------------------------------------------------------------
# deposit                                                         ParticleMesh.jl:289-407
# --------------------------------------------------------------------------------------


def cicdensity(rho, x, m):
    """CIC scatter, summation order identical to the reference loop. :355-407"""
    n3, n2, n1 = rho.shape
    rank = x.shape[1]
    flat = rho.reshape(-1)
    if rank == 2:
        i, ip, dx = _cell(x[:, 0], n1)
        j, jp, dy = _cell(x[:, 1], n2)
        # order :376-379  (ii,ij) (ii,jp1) (ip1,ij) (ip1,jp1)
        idx = np.stack([i + n1 * j, i + n1 * jp, ip + n1 * j, ip + n1 * jp], axis=1)
        w = np.stack([m * (1.0 - dx) * (1.0 - dy), m * (1.0 - dx) * dy,
                      m * dx * (1.0 - dy), m * dx * dy], axis=1)
    elif rank == 3:
        i, ip, dx = _cell(x[:, 0], n1)
        j, jp, dy = _cell(x[:, 1], n2)
        k, kp, dz = _cell(x[:, 2], n3)

        def lin(a, b, c):
            return a + n1 * (b + n2 * c)
        # order :395-402
        idx = np.stack([lin(i, j, k), lin(i, jp, k), lin(ip, j, k), lin(ip, jp, k),
                        lin(i, j, kp), lin(i, jp, kp), lin(ip, j, kp), lin(ip, jp, kp)], axis=1)
        w = np.stack([m * (1 - dx) * (1 - dy) * (1 - dz), m * (1 - dx) * dy * (1 - dz),
                      m * dx * (1 - dy) * (1 - dz), m * dx * dy * (1 - dz),
                      m * (1 - dx) * (1 - dy) * dz, m * (1 - dx) * dy * dz,
                      m * dx * (1 - dy) * dz, m * dx * dy * dz], axis=1)
    else:
        raise ValueError("rank must be 2 or 3")
    # np.add.at is unbuffered and sequential: particle-major, corner-minor == reference order
    np.add.at(flat, idx.reshape(-1), w.reshape(-1))
    return None


def ngpdensity(rho, x, m):
    """NGP scatter. :326-353 (rank 2: floor(x*N)+1 ; rank 3: floor(x*N+1.))."""
    n3, n2, n1 = rho.shape
    rank = x.shape[1]
    flat = rho.reshape(-1)
    if rank == 2:
        i = np.mod(np.floor(x[:, 0] * float(n1)).astype(np.int64), n1)
        j = np.mod(np.floor(x[:, 1] * float(n2)).astype(np.int64), n2)
        idx = i + n1 * j
    else:
        i = np.mod(np.floor(x[:, 0] * float(n1) + 1.0).astype(np.int64) - 1, n1)
        j = np.mod(np.floor(x[:, 1] * float(n2) + 1.0).astype(np.int64) - 1, n2)
        k = np.mod(np.floor(x[:, 2] * float(n3) + 1.0).astype(np.int64) - 1, n3)
        idx = i + n1 * (j + n2 * k)
    np.add.at(flat, idx, np.full(idx.shape, m, dtype=np.float64))
    return None


def deposit(rho, x, m, interpolation="none"):
    """Zero fill then scatter. :289-303"""
    rho[...] = 0.0
    if interpolation == "none":
        ngpdensity(rho, x, m)
    elif interpolation == "cic":
        cicdensity(rho, x, m)
    else:
        raise ValueError("interpolation must be 'none' or 'cic' ('sic' is dead code, :298)")
    return None


# --------------------------------------------------------------------------------------
# potential                                                       ParticleMesh.jl:417-505
# --------------------------------------------------------------------------------------


def _rfft(a):
    # Julia Base rfft: all dims, dim 1 (fastest) halved == numpy rfftn over all axes (last halved)
    return sfft.rfftn(a, workers=1)


def _irfft(rt, shape):
    return sfft.irfftn(rt, s=shape, workers=1)


def compute_potential(phi, rho, rt):
    """rt = rfft(rho); rt /= Ginv; rt[0]=0; phi = irfft(rt)/Nx^2. :417-421, 443-505"""
    n3, n2, n1 = rho.shape
    rt[...] = _rfft(rho)
    ii = np.arange(rt.shape[2])
    jj = np.arange(n2)
    kk = np.arange(n3)
    kx = 2.0 * math.pi * ii / n1
    ky = 2.0 * math.pi * jj / n2
    # :453-455 / :489-491  (wrap above N/2+1 -- only sin^2 is used, so it is immaterial)
    ngl2y = int(round(n2 / 2)) if n3 > 1 else n2 // 2
    ky = np.where(jj + 1 > ngl2y + 1, 2.0 * math.pi * (jj - n2) / n2, ky)
    if n3 > 1:
        kz = 2.0 * math.pi * kk / n3
        ngl2z = int(round(n3 / 2))
        kz = np.where(kk + 1 > ngl2z + 1, 2.0 * math.pi * (kk - n3) / n3, kz)
        ginv = -4.0 * ((np.sin(kx / 2) ** 2)[None, None, :] + (np.sin(ky / 2) ** 2)[None, :, None]
                       + (np.sin(kz / 2) ** 2)[:, None, None])
    else:
        ginv = -4.0 * ((np.sin(kx / 2) ** 2)[None, None, :] + (np.sin(ky / 2) ** 2)[None, :, None])
    nz = ginv != 0.0
    rt[nz] = rt[nz] / ginv[nz]
    rt[0, 0, 0] = 0.0
    # ".* 1/Nglx^2" parses as (A .* 1) / Nglx^2 ; uses Nx for every axis (:466,:502)
    phi[...] = _irfft(rt, rho.shape) * 1 / float(n1) ** 2
    return None


# --------------------------------------------------------------------------------------
# gradient / acceleration                                         ParticleMesh.jl:507-573, 631-637
# --------------------------------------------------------------------------------------


def deriv1(xd, x):
    """xd[...,c] = Ng1 * (x[+1 along c] - x[-1 along c]) / 2 (fac = Ng1 for every axis, :568)."""
    n3, n2, n1 = x.shape
    xd[..., 0] = (np.roll(x, -1, axis=2) - np.roll(x, 1, axis=2)) / 2
    xd[..., 1] = (np.roll(x, -1, axis=1) - np.roll(x, 1, axis=1)) / 2
    if n3 > 1:
        xd[..., 2] = (np.roll(x, -1, axis=0) - np.roll(x, 1, axis=0)) / 2
    # 2-D: third component is never written and stays as allocated (zeros, :754)
    xd *= float(n1)
    return None


def compute_acceleration(a, phi):
    """a = -grad(phi). :631-637"""
    deriv1(a, phi)
    np.negative(a, out=a)
    return None


# --------------------------------------------------------------------------------------
# back-interpolation                                              ParticleMesh.jl:640-721
# --------------------------------------------------------------------------------------


def interpVecToPoints(pa, a, x, interpolation="none", bugcompat=False):
    """Gather the vector field a at the particles.

    bugcompat=True reproduces :706 / :647 literally (k = floor(x3)*Ng3 + 1) for x3 in [0,1)
    -- documented in SURVEY.md F7; parity runs use the intended index (bugcompat=False).
    """
    n3, n2, n1, _ = a.shape
    rank = x.shape[1]
    if interpolation == "none":
        i = np.mod(np.floor(x[:, 0] * float(n1)).astype(np.int64), n1)
        j = np.mod(np.floor(x[:, 1] * float(n2)).astype(np.int64), n2)
        if rank == 3:
            if bugcompat:
                k = np.mod(np.floor(x[:, 2]).astype(np.int64) * n3, n3)
            else:
                k = np.mod(np.floor(x[:, 2] * float(n3)).astype(np.int64), n3)
            pa[:, :] = a[k, j, i, :3]
        else:
            pa[:, :] = a[0, j, i, :2]
        return None
    if interpolation != "cic":
        raise ValueError("interpolation must be 'none' or 'cic'")
    i, ip, dx = _cell(x[:, 0], n1)
    j, jp, dy = _cell(x[:, 1], n2)
    if rank == 2:
        for mm in range(2):
            am = a[0, :, :, mm]
            pa[:, mm] = (am[j, i] * (1.0 - dx) * (1.0 - dy) + am[j, ip] * dx * (1.0 - dy)
                         + am[jp, i] * (1.0 - dx) * dy + am[jp, ip] * dx * dy)
        return None
    if bugcompat:
        k1 = np.floor(x[:, 2]).astype(np.int64) * n3 + 1      # literal :706 (1-based)
        dz = (x[:, 2] * float(n3) - k1.astype(np.float64)) + 1.0
        k = np.mod(k1 - 1, n3)
        kp = np.where(k1 == n3, 0, np.mod(k1, n3))
    else:
        k, kp, dz = _cell(x[:, 2], n3)
    for mm in range(3):
        am = a[..., mm]
        pa[:, mm] = (am[k, j, i] * (1.0 - dx) * (1.0 - dy) * (1.0 - dz)
                     + am[k, j, ip] * dx * (1.0 - dy) * (1.0 - dz)
                     + am[k, jp, i] * (1.0 - dx) * dy * (1.0 - dz)
                     + am[k, jp, ip] * dx * dy * (1.0 - dz)
                     + am[kp, j, i] * (1.0 - dx) * (1.0 - dy) * dz
                     + am[kp, j, ip] * dx * (1.0 - dy) * dz
                     + am[kp, jp, i] * (1.0 - dx) * dy * dz
                     + am[kp, jp, ip] * dx * dy * dz)
    return None


# --------------------------------------------------------------------------------------
# drift / kick                                                    ParticleMesh.jl:723-736, 800-817
# --------------------------------------------------------------------------------------


def drift(x, v, dt):
    """x += dt*v. :723-728"""
    x[...] = x + dt * v
    return None


def kick(v, a, dt, ascale=None, dadt=None):
    """Static v += dt*a (:731-736) or comoving semi-implicit kick (:800-807)."""
    if ascale is None:
        v[...] = v + dt * a
        return None
    vmid = v + a * dt / 2                      # "v .+ pa .* dt/2" == v + (pa*dt)/2
    v += (-vmid * (dadt / ascale) + a / ascale) * dt
    return None


# --------------------------------------------------------------------------------------
# cosmology scalars                                               cosmology.jl:4-46, 61-72
# --------------------------------------------------------------------------------------

gyr_in_s = 1.0e9 * 31_556_952                  # cosmology.jl:4
# Cosmology.jl (unpinned, not vendored): hubble_time_gyr0(c) = 9.77813952 / c.h in the releases of the reference's
# era, quoted from memory.  Corroborated by the reference's own constants: da/dt = dadt(a) (cosmology.jl:30-37) holds
# for a(t) = a_from_t_internal(t) only if t_H = 2.519445e17 / gyr_in_s / sqrt(2/3) = 9.778122 Gyr (cosmology.jl:4, :69);
# 9.77813952 matches that to 1.8e-6 (tests/test_oracle_kats.py).  Rounds 1-2 used 9.77793 (2e-5 off).
HUBBLE_TIME_GYR = 9.77813952


class Cosmo:
    """Restates Cosmology.jl's `cosmology(h=, OmegaM=, OmegaK=0, OmegaR=nothing)` constructor
    and `age_gyr` (published algorithm: age = t_H * int_0^a x dx / sqrt(Or + Om x + Ok x^2 + OL x^4))."""

    def __init__(self, h=0.69, OmegaM=0.29, OmegaK=0.0, OmegaR=None, Tcmb=2.7255, Neff=3.04, hubble_time_gyr=None):
        self.tH = float(HUBBLE_TIME_GYR if hubble_time_gyr is None else hubble_time_gyr)
        if OmegaR is None:
            og = 4.48131e-7 * Tcmb ** 4 / h ** 2
            on = Neff * og * (7.0 / 8.0) * (4.0 / 11.0) ** (4.0 / 3.0)
            OmegaR = og + on
        self.h = float(h)
        self.Om = float(OmegaM)
        self.Ok = float(OmegaK)
        self.Or = float(OmegaR)
        self.OL = 1.0 - self.Ok - self.Om - self.Or

    def a2E(self, a):
        return math.sqrt(self.Or + self.Om * a + self.Ok * a * a + self.OL * a ** 4)

    def age_gyr(self, z):
        a1 = 1.0 / (1.0 + z)
        # substitution a = s^2 removes the sqrt(a) endpoint behaviour when Or == 0
        val, _ = integrate.quad(lambda s: 2.0 * s ** 3 / self.a2E(s * s), 0.0, math.sqrt(a1),
                                epsabs=0.0, epsrel=1e-13, limit=200)
        return self.tH / self.h * val


def get_units_T(cosmo, zinit):
    """u.T of cosmology.jl:69"""
    return 2.519445e17 / math.sqrt(cosmo.Om) / cosmo.h / (1 + zinit) ** 1.5


def z_from_t(cosmo, t_gyr):
    """cosmology.jl:7-11 (fzero bracket [-1+1e-15, 1e80] -> brentq on a practical bracket)."""
    return optimize.brentq(lambda z: cosmo.age_gyr(z) - t_gyr, -1.0 + 1e-9, 1e8,
                           xtol=1e-15, rtol=8.9e-16, maxiter=500)


def a_from_t_internal(cosmo, t, zinit, zwherea1=0.0):
    """cosmology.jl:19-23 + :13-17"""
    time = t * (get_units_T(cosmo, zinit) / gyr_in_s)
    z = z_from_t(cosmo, time)
    return (1.0 + zwherea1) / (1.0 + z)


def dadt(cosmo, a, zwherea1=0.0):
    """cosmology.jl:30-37 (Peebles93 eq. 13.3 in Enzo units)"""
    tv = a / (1.0 + zwherea1)
    ok = 1 - cosmo.Om - cosmo.OL - cosmo.Or
    return math.sqrt(2.0 / (3.0 * cosmo.Om * a) * (cosmo.Om + ok * tv + cosmo.OL * tv ** 3))


def dadt_from_t_internal(cosmo, t, zinit):
    """cosmology.jl:39-42"""
    a = a_from_t_internal(cosmo, t, zinit, zwherea1=zinit)
    return dadt(cosmo, a, zwherea1=zinit)


def z_from_t_internal(cosmo, t, zinit):
    """cosmology.jl:25-28 (a with zwherea1=0, then z_from_a with zwherea1=0)"""
    a = a_from_t_internal(cosmo, t, zinit)
    return 1.0 / a - 1.0


def start_stop_times(cosmo, zstart, zend):
    """ParticleMesh.jl:129-138"""
    ut = get_units_T(cosmo, zstart)
    return cosmo.age_gyr(zstart) * (gyr_in_s / ut), cosmo.age_gyr(zend) * (gyr_in_s / ut)


# --------------------------------------------------------------------------------------
# initial conditions                                              ParticleMesh.jl:161-198, 257-287
# --------------------------------------------------------------------------------------


def initialize_particles_uniform(pdims, rank):
    """Lattice (i-1/2)/Npd, i fastest. :257-287"""
    p1, p2, p3 = (list(pdims) + [1, 1])[:3]
    if rank == 2:
        jj, ii = np.meshgrid(np.arange(1, p2 + 1), np.arange(1, p1 + 1), indexing="ij")
        x = np.stack([ii.ravel(), jj.ravel()], axis=1).astype(np.float64)
        x = (x - 0.5) / np.array([p1, p2], dtype=np.float64)
    else:
        kk, jj, ii = np.meshgrid(np.arange(1, p3 + 1), np.arange(1, p2 + 1), np.arange(1, p1 + 1),
                                 indexing="ij")
        x = np.stack([ii.ravel(), jj.ravel(), kk.ravel()], axis=1).astype(np.float64)
        x = (x - 0.5) / np.array([p1, p2, p3], dtype=np.float64)
    return x


def zeldovich_pancake(pdims, rank, zc, zinit, a, adot):
    """ZeldovichPancakeParticles. :177-198 (kZ = 1)."""
    kz_ = 1.0
    A = 5.0 * (1.0 + zc) / (2 * kz_) * 1.0 / (1.0 + zinit)
    x = initialize_particles_uniform(pdims, rank)
    v = np.zeros_like(x)
    v[:, 0] = 2.0 / 5 * A * adot * np.sin(2 * math.pi * kz_ * x[:, 0])
    x[:, 0] += 2.0 / 5 * A * a * np.sin(2 * math.pi * kz_ * x[:, 0])
    make_periodic(x)
    return x, v


def synthetic_modes(seed, nmodes, kmax):
    """Seeded plane-wave set for the synthetic Zel'dovich ICs used by bench/tests (NOT in the
    reference, whose P(k) ICs are unseeded and broken -- SURVEY.md F8).  splitmix64 stream,
    mirrored bit for bit by the device generator (noname_b200/csrc/ics.cu)."""
    state = np.uint64(seed)
    out = []

    def nxt():
        nonlocal state
        with np.errstate(over="ignore"):
            state = state + np.uint64(0x9E3779B97F4A7C15)
            z = state
            z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
            z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
            z = z ^ (z >> np.uint64(31))
        return int(z)

    while len(out) < nmodes:
        kx = nxt() % (2 * kmax + 1) - kmax
        ky = nxt() % (2 * kmax + 1) - kmax
        kz = nxt() % (2 * kmax + 1) - kmax
        ph = (nxt() >> 11) * (1.0 / 9007199254740992.0)
        if kx == 0 and ky == 0 and kz == 0:
            continue
        out.append((kx, ky, kz, ph))
    return out


def synthetic_zeldovich(pdims, rank, seed, nmodes, kmax, disp_rms, vfac):
    """x = q + psi, v = vfac*psi with psi = sum_m A * khat_m/|k_m| * sin(2pi(k_m.q + phase_m)),
    scaled so the analytic rms displacement per axis is disp_rms (box units)."""
    q = initialize_particles_uniform(pdims, rank)
    modes = synthetic_modes(seed, nmodes, kmax)
    if rank == 2:
        modes = [(kx, ky, 0, ph) for (kx, ky, kz, ph) in modes if (kx, ky) != (0, 0)]
    psi = np.zeros_like(q)
    norm2 = 0.0
    for (kx, ky, kz, ph) in modes:
        k2 = float(kx * kx + ky * ky + kz * kz)
        norm2 += 0.5 / k2
    amp = disp_rms * math.sqrt(rank) / math.sqrt(norm2)
    for (kx, ky, kz, ph) in modes:
        k2 = float(kx * kx + ky * ky + kz * kz)
        arg = kx * q[:, 0] + ky * q[:, 1] + (kz * q[:, 2] if rank == 3 else 0.0) + ph
        s = np.sin(2.0 * math.pi * arg)
        psi[:, 0] += (amp * kx / k2) * s
        psi[:, 1] += (amp * ky / k2) * s
        if rank == 3:
            psi[:, 2] += (amp * kz / k2) * s
    x = q + psi
    v = vfac * psi
    make_periodic(x)
    return x, v


def _mix64(z):
    z = np.asarray(z, dtype=np.uint64)
    with np.errstate(over="ignore"):
        z = z + np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return z ^ (z >> np.uint64(31))


def _u01(h):
    return ((h >> np.uint64(11)).astype(np.float64) + 0.5) * (1.0 / 9007199254740992.0)


def synthetic_clustered(pdims, rank, seed, nblobs, sigma, vdisp):
    """Clustered stand-in for BASELINE config 4 (NOT in the reference): mirrors the device generator
    k_init_clustered hash for hash.  Half the particles on the lattice, half in Gaussian blobs with a
    b^-1/2 mass function."""
    x = initialize_particles_uniform(pdims, rank)
    v = np.zeros_like(x)
    n = x.shape[0]
    with np.errstate(over="ignore"):
        h0 = _mix64(np.uint64(seed) ^ _mix64(np.arange(n, dtype=np.uint64)))
        cl = (h0 & np.uint64(1)) == np.uint64(1)
        u = _u01(_mix64(h0 + np.uint64(1)))
        sq = 1.0 + u * (math.sqrt(float(nblobs)) - 1.0)
        b = np.minimum((sq * sq).astype(np.int64), nblobs).astype(np.uint64)
        hb = _mix64(np.uint64(seed) * np.uint64(0x100000001B3) + b)
        for d in range(rank):
            c = _u01(_mix64(hb + np.uint64(11 * (d + 1))))
            u1 = _u01(_mix64(h0 + np.uint64(101 * (d + 1))))
            u2 = _u01(_mix64(h0 + np.uint64(211 * (d + 1))))
            r = np.sqrt(-2.0 * np.log(u1))
            xd = c + sigma * r * np.cos(2.0 * math.pi * u2)
            xd -= np.floor(xd)
            x[cl, d] = xd[cl]
            v[cl, d] = (vdisp * r * np.sin(2.0 * math.pi * u2))[cl]
    return x, v


def grf_spectrum(pdims, comp, seed, ps_index, peaks=None):
    """Half spectrum c_d[k, j, i] (i <= P1/2) of displacement component `comp`, PowerSpectrumParticles :224-240 with
    scaleFreePS P(k) = k^n (InputPowerSpectra.jl:26; the norm drops out of the rescaled field): amplitude
    sqrt(P(|k|)) (g1 - i g2)/2 / k^2 * k_d, k in rad/cell via fftfreq (:202-206), k = 0 skipped.  Deviations SURVEY.md
    8(d) prescribes and nnpm_init_grf implements: SEEDED draws (mix64 keyed by the mode's linear index), one complex
    amplitude per mode shared by the components, Hermitian i = 0 and i = P1/2 planes."""
    P1, P2, P3 = (int(q) for q in pdims)
    icn = P1 // 2 + 1
    K, J, I = np.meshgrid(np.arange(P3), np.arange(P2), np.arange(icn), indexing="ij")

    def kf(idx, n):
        w = np.where(idx > n // 2, idx - n, idx).astype(np.float64)
        return (2.0 * math.pi * w) / float(n)

    edge = (I == 0) | (2 * I == P1)
    Jm, Km = (P2 - J) % P2, (P3 - K) % P3
    a, b = J + P2 * K, Jm + P2 * Km
    conj = edge & (b < a)
    real_only = edge & (b == a)
    Jc, Kc = np.where(conj, Jm, J), np.where(conj, Km, K)
    kk = [kf(I, P1), kf(Jc, P2), kf(Kc, P3) if P3 > 1 else np.zeros(I.shape)]
    k2 = kk[0] * kk[0] + kk[1] * kk[1] + kk[2] * kk[2]
    nz = k2 > 0.0
    k2s = np.where(nz, k2, 1.0)
    ka = np.sqrt(k2s)
    psv = np.sqrt(np.power(ka, ps_index))
    if peaks is not None:
        # multiplePeaksPS, InputPowerSpectra.jl:28-36: P = P_scalefree(k) if |k - k_p| / k < lnwidth_p / 2 for some p, else 0
        inside = np.zeros(ka.shape, dtype=bool)
        for kp, w in zip(*peaks):
            inside |= np.abs(ka - float(kp)) / ka < float(w) / 2
        nz = nz & inside
    with np.errstate(over="ignore"):
        lin = I.astype(np.uint64) + np.uint64(icn) * (Jc.astype(np.uint64) + np.uint64(P2) * Kc.astype(np.uint64))
        h = _mix64(np.uint64(seed) ^ _mix64(lin))
        u1, u2 = _u01(_mix64(h + np.uint64(1))), _u01(_mix64(h + np.uint64(2)))
    r = np.sqrt(-2.0 * np.log(u1))
    g1, g2 = r * np.cos(2.0 * math.pi * u2), r * np.sin(2.0 * math.pi * u2)
    re = np.where(nz, psv * g1 / k2s / 2.0, 0.0) * kk[comp]
    im = np.where(nz, -(psv * g2 / k2s) / 2.0, 0.0) * kk[comp]
    im = np.where(conj, -im, im)
    im = np.where(real_only, 0.0, im)
    return re + 1j * im


def synthetic_grf(pdims, rank, seed, ps_index, disp_rms, vfac, peaks=None):
    """Seeded Gaussian-random-field Zel'dovich ICs (checker of nnpm_init_grf): x = wrap(q + s psi), v = vfac s psi with
    psi_d = irfft(c_d) on the particle lattice (:241) and s rescaling psi to a per-component RMS of disp_rms."""
    P1, P2, P3 = (int(q) for q in pdims)
    if rank == 2:
        P3 = 1
    q = initialize_particles_uniform(pdims, rank)
    psi = np.empty_like(q)
    for d in range(rank):
        c = grf_spectrum((P1, P2, P3), d, seed, ps_index, peaks)
        f = sfft.irfftn(c, s=(P3, P2, P1), axes=(0, 1, 2)) * float(P1 * P2 * P3)   # unnormalised, like the device
        psi[:, d] = f.reshape(-1)
    rms = math.sqrt(float(np.sum(psi * psi)) / (psi.shape[0] * rank))
    s = disp_rms / rms if rms > 0 else 0.0
    psi *= s
    x = q + psi
    v = vfac * psi
    make_periodic(x)
    return x, v


# --------------------------------------------------------------------------------------
# P(k)                                                            ParticleMesh.jl:65-100, 202-206
# --------------------------------------------------------------------------------------


def fftfreq(n):
    """k = 2pi (i - N*(i > N/2)) / N for 0-based i (1-based ijk > s+1). :202-206"""
    i = np.arange(n)
    s = n // 2
    return 2.0 * math.pi * (i - (i + 1 > s + 1) * n) / n


def powerSpectrum(x, m, dims, interpolation="cic", box=(1.0, 1.0, 1.0)):
    """:65-100.  dims = (N1,N2,N3). Returns karray, power."""
    n1, n2, n3 = dims
    make_periodic(x)
    rho = np.zeros((n3, n2, n1))
    deposit(rho, x, m, interpolation=interpolation)
    delta = rho / rho.mean() - 1
    deltak2 = np.abs(sfft.fftn(delta, workers=1)) ** 2 * float(np.prod(box)) / float(n1 * n2 * n3) ** 2
    nmax = max(dims)
    dk = 2 * math.pi / nmax
    lenk = int(round(nmax / 2))
    karray = np.arange(1, lenk + 1) * dk
    kxv, kyv, kzv = fftfreq(n1), fftfreq(n2), fftfreq(n3)
    ka = np.sqrt(kxv[None, None, :] ** 2 + kyv[None, :, None] ** 2 + kzv[:, None, None] ** 2)
    nb = np.rint(ka / dk).astype(np.int64)        # Julia round: half away from even? see note
    ok = (nb >= 1) & (nb <= lenk)
    power = np.bincount(nb[ok] - 1, weights=deltak2[ok], minlength=lenk)
    counts = np.bincount(nb[ok] - 1, minlength=lenk).astype(np.float64)
    with np.errstate(invalid="ignore", divide="ignore"):
        power = power / counts
    return karray, power


# --------------------------------------------------------------------------------------
# step bodies and evolve loops                                    ParticleMesh.jl:738-796, 820-894
# --------------------------------------------------------------------------------------


class Fields:
    """The per-loop buffers the reference allocates once (:754-757, :833-838)."""

    def __init__(self, dims, npart, rank):
        n1, n2, n3 = dims
        self.rho = np.ones((n3, n2, n1))
        self.phi = np.ones((n3, n2, n1))
        self.acc = np.zeros((n3, n2, n1, 3))
        self.pa = np.zeros((npart, rank))
        self.rt = np.zeros((n3, n2, n1 // 2 + 1), dtype=np.complex128)


def step_static(f, x, v, m, dt, dep="cic", back="cic", bugcompat=False):
    """Body of evolvePM's loop, :766-778."""
    make_periodic(x)
    deposit(f.rho, x, m, interpolation=dep)
    rho = f.rho - f.rho.mean()
    compute_potential(f.phi, rho, f.rt)
    compute_acceleration(f.acc, f.phi)
    interpVecToPoints(f.pa, f.acc, x, interpolation=back, bugcompat=bugcompat)
    drift(x, v, dt / 2)
    kick(v, f.pa, dt)
    drift(x, v, dt / 2)


def step_comoving(f, x, v, m, dt, a_d1, a_k, adot_k, a_d2, dep="cic", back="cic", bugcompat=False):
    """Body of evolvePMCosmology's loop, :853-868, with the four host scalars passed in."""
    drift(x, v, dt / 2 / a_d1)
    make_periodic(x)
    deposit(f.rho, x, m, interpolation=dep)
    compute_potential(f.phi, f.rho - f.rho.mean(), f.rt)
    compute_acceleration(f.acc, f.phi)
    interpVecToPoints(f.pa, f.acc, x, interpolation=back, bugcompat=bugcompat)
    kick(v, f.pa, dt, a_k, adot_k)
    drift(x, v, dt / 2 / a_d2)
    make_periodic(x)


def evolvePM(x, v, m, dims, conf, dep="cic", back="cic", bugcompat=False, log=None):
    """evolvePM time loop incl. dt control, :759-789. conf: StartTime, StopTime, StartCycle,
    StopCycle, InitialDt, CourantFactor, MaximumDt. Returns (time, cycle, fields)."""
    f = Fields(dims, x.shape[0], x.shape[1])
    dx = 1.0 / max(dims)
    dt = conf["InitialDt"]
    t = conf["StartTime"]
    cyc = conf.get("StartCycle", 0)
    while t < conf["StopTime"] and cyc < conf.get("StopCycle", 1000000):
        step_static(f, x, v, m, dt, dep, back, bugcompat)
        t += dt
        if log is not None:
            log.append((t, dt, float(np.max(f.pa)), float(np.max(v))))
        dt = conf["CourantFactor"] * dx / np.max(np.abs(v) + 1e-30)
        dt = min(dt, conf["MaximumDt"])
        if t + dt > conf["StopTime"]:
            dt = (1.0 + 1e-15) * (conf["StopTime"] - t)
        cyc += 1
    return t, cyc, f


def evolvePMCosmology(x, v, m, dims, conf, cosmo, dep="cic", back="cic", bugcompat=False, log=None):
    """evolvePMCosmology time loop, :840-886. conf additionally: InitialRedshift,
    DtFractionOfTime."""
    f = Fields(dims, x.shape[0], x.shape[1])
    dx = 1.0 / max(dims)
    dt = conf["InitialDt"]
    zinit = conf["InitialRedshift"]
    t = conf["StartTime"]
    cyc = conf.get("StartCycle", 0)
    while t < conf["StopTime"] and cyc < conf.get("StopCycle", 1000000):
        a1 = a_from_t_internal(cosmo, t + dt / 2 / 2, zinit, zwherea1=zinit)
        ak = a_from_t_internal(cosmo, t + dt / 2, zinit, zwherea1=zinit)
        adk = dadt_from_t_internal(cosmo, t + dt / 2, zinit)
        a2 = a_from_t_internal(cosmo, t + 3.0 * dt / 2 / 2, zinit, zwherea1=zinit)
        step_comoving(f, x, v, m, dt, a1, ak, adk, a2, dep, back, bugcompat)
        t += dt
        if log is not None:
            log.append((t, dt, float(np.max(f.pa)), float(np.max(v))))
        dt = conf["CourantFactor"] * dx / np.max(np.abs(v) + 1e-30)
        dt = min(dt, conf["DtFractionOfTime"] * t, conf["MaximumDt"])
        if t + dt > conf["StopTime"]:
            dt = (1.0 + 1e-6) * (conf["StopTime"] - t)
        cyc += 1
    return t, cyc, f


# --------------------------------------------------------------------------------------
# literal scalar-loop transcriptions (tiny inputs only) -- pin the vectorised code above
# --------------------------------------------------------------------------------------


def cicdensity_loops(rho, x, m):
    """Line-by-line :383-405 / :367-382 with 0-based indices."""
    n3, n2, n1 = rho.shape
    rank = x.shape[1]
    for n in range(x.shape[0]):
        t = x[n, 0] * float(n1)
        ii = math.floor(t) + 1
        dx = t - ii + 1
        ii = (ii - 1) % n1 + 1
        ip1 = 1 if ii == n1 else ii + 1
        t = x[n, 1] * float(n2)
        ij = math.floor(t) + 1
        dy = t - ij + 1
        ij = (ij - 1) % n2 + 1
        jp1 = 1 if ij == n2 else ij + 1
        if rank == 2:
            rho[0, ij - 1, ii - 1] += m * (1. - dx) * (1. - dy)
            rho[0, jp1 - 1, ii - 1] += m * (1. - dx) * dy
            rho[0, ij - 1, ip1 - 1] += m * dx * (1. - dy)
            rho[0, jp1 - 1, ip1 - 1] += m * dx * dy
            continue
        t = x[n, 2] * float(n3)
        ik = math.floor(t) + 1
        dz = t - ik + 1
        ik = (ik - 1) % n3 + 1
        kp1 = 1 if ik == n3 else ik + 1
        rho[ik - 1, ij - 1, ii - 1] += m * (1 - dx) * (1 - dy) * (1 - dz)
        rho[ik - 1, jp1 - 1, ii - 1] += m * (1 - dx) * dy * (1 - dz)
        rho[ik - 1, ij - 1, ip1 - 1] += m * dx * (1 - dy) * (1 - dz)
        rho[ik - 1, jp1 - 1, ip1 - 1] += m * dx * dy * (1 - dz)
        rho[kp1 - 1, ij - 1, ii - 1] += m * (1 - dx) * (1 - dy) * dz
        rho[kp1 - 1, jp1 - 1, ii - 1] += m * (1 - dx) * dy * dz
        rho[kp1 - 1, ij - 1, ip1 - 1] += m * dx * (1 - dy) * dz
        rho[kp1 - 1, jp1 - 1, ip1 - 1] += m * dx * dy * dz


def potential_loops(phi, rho, rt):
    """Line-by-line :473-505 / :443-469 (per-element sin, as the reference)."""
    n3, n2, n1 = rho.shape
    rt[...] = _rfft(rho)
    for k in range(n3):
        kz = 2 * math.pi * k / n3
        for j in range(n2):
            ky = 2 * math.pi * j / n2
            for i in range(rt.shape[2]):
                kx = 2 * math.pi * i / n1
                if n3 > 1:
                    ginv = -4. * (math.sin(kx / 2) ** 2 + math.sin(ky / 2) ** 2 + math.sin(kz / 2) ** 2)
                else:
                    ginv = -4. * (math.sin(kx / 2) ** 2 + math.sin(ky / 2) ** 2)
                if not ginv == 0.0:
                    rt[k, j, i] /= ginv
    rt[0, 0, 0] = 0.
    phi[...] = _irfft(rt, rho.shape) * 1 / float(n1) ** 2


def deriv1_loops(xd, x):
    """Line-by-line :549-573 / :528-546."""
    n3, n2, n1 = x.shape
    for k in range(n3):
        kp1 = 0 if k == n3 - 1 else k + 1
        km1 = n3 - 1 if k == 0 else k - 1
        for j in range(n2):
            jp1 = 0 if j == n2 - 1 else j + 1
            jm1 = n2 - 1 if j == 0 else j - 1
            for i in range(n1):
                ip1 = 0 if i == n1 - 1 else i + 1
                im1 = n1 - 1 if i == 0 else i - 1
                xd[k, j, i, 0] = (x[k, j, ip1] - x[k, j, im1]) / 2
                xd[k, j, i, 1] = (x[k, jp1, i] - x[k, jm1, i]) / 2
                if n3 > 1:
                    xd[k, j, i, 2] = (x[kp1, j, i] - x[km1, j, i]) / 2
    xd *= float(n1)


def cic_interp_loops(pa, a, x):
    """Line-by-line :694-721 / :676-692 with the intended k index (F7 fixed)."""
    n3, n2, n1, _ = a.shape
    rank = x.shape[1]
    for n in range(x.shape[0]):
        t = x[n, 0] * float(n1)
        i = math.floor(t) + 1
        dx = t - i + 1
        i = (i - 1) % n1
        ip1 = 0 if i == n1 - 1 else i + 1
        t = x[n, 1] * float(n2)
        j = math.floor(t) + 1
        dy = t - j + 1
        j = (j - 1) % n2
        jp1 = 0 if j == n2 - 1 else j + 1
        if rank == 2:
            for mm in range(2):
                pa[n, mm] = (a[0, j, i, mm] * (1. - dx) * (1. - dy) + a[0, j, ip1, mm] * dx * (1. - dy)
                             + a[0, jp1, i, mm] * (1. - dx) * dy + a[0, jp1, ip1, mm] * dx * dy)
            continue
        t = x[n, 2] * float(n3)
        k = math.floor(t) + 1
        dz = t - k + 1
        k = (k - 1) % n3
        kp1 = 0 if k == n3 - 1 else k + 1
        for mm in range(3):
            pa[n, mm] = (a[k, j, i, mm] * (1. - dx) * (1. - dy) * (1. - dz)
                         + a[k, j, ip1, mm] * dx * (1. - dy) * (1. - dz)
                         + a[k, jp1, i, mm] * (1. - dx) * dy * (1. - dz)
                         + a[k, jp1, ip1, mm] * dx * dy * (1. - dz)
                         + a[kp1, j, i, mm] * (1. - dx) * (1. - dy) * dz
                         + a[kp1, j, ip1, mm] * dx * (1. - dy) * dz
                         + a[kp1, jp1, i, mm] * (1. - dx) * dy * dz
                         + a[kp1, jp1, ip1, mm] * dx * dy * dz)
