"""Host-side cosmology scalars for the comoving step -- the product twin of
src/cosmology.jl:4-46, 61-72 of the reference (a(t), da/dt, z(t), Enzo code units).

The reference root-finds Cosmology.jl's `age_gyr` with Roots.jl's `fzero` five times per step
(cosmology.jl:7-11); both packages are un-vendored and unpinned.  Here age(a) is a 1-D integral
  age = t_H * int_0^a x dx / sqrt(Or + Om x + Ok x^2 + OL x^4),   t_H = 9.77813952 Gyr / h
evaluated by composite Gauss-Legendre in s = sqrt(x) (closed form for Or = Ok = 0), and a(t) is
found by safeguarded Newton iterations on it (dt/da is the integrand) -- microseconds per call,
so the five scalars per step stay off the GPU step's critical path.
"""
from __future__ import annotations

import math
import re

import numpy as np

gyr_in_s = 1.0e9 * 31_556_952            # cosmology.jl:4
# Cosmology.jl (un-vendored, unpinned): hubble_time_gyr0(c) = 9.77813952 / c.h in the releases of the reference's era
# (quoted from memory of the package -- "parity unpinned").  The value is corroborated by the reference's own constants:
# da/dt = dadt(a) (cosmology.jl:30-37) holds for a(t) = a_from_t_internal(t) only if
# t_H = 2.519445e17 / gyr_in_s / sqrt(2/3) = 9.778122 Gyr (get_units' T, cosmology.jl:4, :69) -- 9.77813952 matches that
# to 1.8e-6, the Julian-year figure 9.7779222 to 2e-5 (tests/test_oracle_kats.py).  Rounds 1-2 of this repo used 9.77793;
# `Cosmology(..., hubble_time_gyr=9.77793)` reproduces them (bench.py does, for its stored P(k) reference).
HUBBLE_TIME_GYR = 9.77813952

_GL_X, _GL_W = np.polynomial.legendre.leggauss(48)


class Cosmology:
    """`cosmology(h=, OmegaM=, OmegaK=0, OmegaR=nothing)` of Cosmology.jl (fields h, Ω_m, Ω_Λ, Ω_r)."""

    def __init__(self, h=0.69, OmegaM=0.29, OmegaK=0.0, OmegaR=None, Tcmb=2.7255, Neff=3.04, hubble_time_gyr=None):
        self.tH = float(HUBBLE_TIME_GYR if hubble_time_gyr is None else hubble_time_gyr)   # Gyr * h
        if OmegaR is None:
            og = 4.48131e-7 * Tcmb ** 4 / h ** 2
            OmegaR = og + Neff * og * (7.0 / 8.0) * (4.0 / 11.0) ** (4.0 / 3.0)
        self.h, self.Om, self.Ok, self.Or = float(h), float(OmegaM), float(OmegaK), float(OmegaR)
        self.OL = 1.0 - self.Ok - self.Om - self.Or

    # dt/da in units of t_H
    def _dtda(self, a):
        return a / math.sqrt(self.Or + self.Om * a + self.Ok * a * a + self.OL * a ** 4)

    def age_th(self, a):
        """age(a) / t_H."""
        if a <= 0.0:
            return 0.0
        if self.Or == 0.0 and self.Ok == 0.0:
            if self.OL == 0.0:
                return 2.0 / 3.0 * a ** 1.5 / math.sqrt(self.Om)
            if self.OL > 0.0:
                return 2.0 / (3.0 * math.sqrt(self.OL)) * math.asinh(math.sqrt(self.OL / self.Om) * a ** 1.5)
        # int_0^a f(x) dx with x = s^2: int_0^sqrt(a) 2 s^3 / sqrt(Or + Om s^2 + Ok s^4 + OL s^8) ds
        sa = math.sqrt(a)
        npan = 8
        edges = np.linspace(0.0, sa, npan + 1)
        half = 0.5 * (edges[1:] - edges[:-1])[:, None]
        mid = 0.5 * (edges[1:] + edges[:-1])[:, None]
        s = mid + half * _GL_X[None, :]
        s2 = s * s
        f = 2.0 * s * s2 / np.sqrt(self.Or + self.Om * s2 + self.Ok * s2 * s2 + self.OL * s2 ** 4)
        return float(np.sum(f * _GL_W[None, :] * half))

    def age_gyr(self, z):
        return self.tH / self.h * self.age_th(1.0 / (1.0 + z))

    def a_of_age_gyr(self, t_gyr):
        """Inverse of age_gyr in terms of a = 1/(1+z): safeguarded Newton."""
        tau = t_gyr * self.h / self.tH
        if tau <= 0.0:
            raise ValueError("time must be positive")
        a = (1.5 * tau * math.sqrt(self.Om)) ** (2.0 / 3.0)      # EdS guess
        lo, hi = 0.0, math.inf
        for _ in range(100):
            f = self.age_th(a) - tau
            if f > 0:
                hi = min(hi, a)
            else:
                lo = max(lo, a)
            step = f / self._dtda(a)
            an = a - step
            if not (lo < an < hi):
                an = 0.5 * (lo + hi) if math.isfinite(hi) else 2.0 * a
            if abs(an - a) <= 4e-16 * abs(an):
                a = an
                break
            a = an
        return a


def parse_cosmology(expr):
    """`Cosmology = cosmology(h=1.,OmegaM=1.0,OmegaR=0.0)` conf value -> Cosmology
    (the reference evals the string, ParticleMesh.jl:128)."""
    if isinstance(expr, Cosmology):
        return expr
    m = re.match(r"\s*cosmology\s*\((.*)\)\s*$", str(expr))
    if not m:
        raise ValueError(f"cannot parse Cosmology = {expr!r}")
    kw = {}
    for part in filter(None, (p.strip() for p in m.group(1).split(","))):
        k, _, val = part.partition("=")
        kw[k.strip()] = float(val)
    allowed = {"h", "OmegaM", "OmegaK", "OmegaR", "Tcmb", "Neff"}
    if set(kw) - allowed:
        raise ValueError(f"unsupported cosmology() arguments {sorted(set(kw) - allowed)}")
    return Cosmology(**kw)


class CodeUnits:
    """get_units, cosmology.jl:61-72 (Enzo unit system)."""

    def __init__(self, cosmo, z, zinit=0.0, boxsize=1.0):
        self.rho = 1.8788e-29 * cosmo.Om * cosmo.h ** 2 * (1.0 + z) ** 3
        self.L = 3.085678e24 * boxsize / cosmo.h / (1.0 + z)
        self.Temp = 1.81723e6 * boxsize ** 2 * cosmo.Om * (1 + zinit)
        self.T = 2.519445e17 / math.sqrt(cosmo.Om) / cosmo.h / (1 + zinit) ** 1.5
        self.V = 1.22475e7 * boxsize * math.sqrt(cosmo.Om) * math.sqrt(1 + zinit)


def get_units(cosmo, z, zinit=0.0, boxsize=1.0):
    return CodeUnits(cosmo, z, zinit=zinit, boxsize=boxsize)


def z_from_t(cosmo, t_gyr):
    """cosmology.jl:7-11"""
    return 1.0 / cosmo.a_of_age_gyr(t_gyr) - 1.0


def a_from_t(cosmo, t_gyr, zwherea1=0.0):
    """cosmology.jl:13-17:  a = (1+zwherea1)/(1+z)"""
    return (1.0 + zwherea1) * cosmo.a_of_age_gyr(t_gyr)


def a_from_t_internal(cosmo, t, zinit, zwherea1=0.0):
    """cosmology.jl:19-23"""
    u = get_units(cosmo, 0.0, zinit=zinit)
    return a_from_t(cosmo, t * (u.T / gyr_in_s), zwherea1=zwherea1)


def z_from_a(cosmo, a, zwherea1=0.0):
    """cosmology.jl:44-46"""
    return (1.0 + zwherea1) / a - 1.0


def z_from_t_internal(cosmo, t, zinit, zwherea1=0.0):
    """cosmology.jl:25-28"""
    return z_from_a(cosmo, a_from_t_internal(cosmo, t, zinit))


def dadt(cosmo, a, zwherea1=0.0):
    """cosmology.jl:30-37 (Peebles93 eq. 13.3)"""
    tv = a / (1.0 + zwherea1)
    ok = 1 - cosmo.Om - cosmo.OL - cosmo.Or
    return math.sqrt(2.0 / (3.0 * cosmo.Om * a) * (cosmo.Om + ok * tv + cosmo.OL * tv ** 3))


def dadt_from_t_internal(cosmo, t, zinit):
    """cosmology.jl:39-42"""
    a = a_from_t_internal(cosmo, t, zinit, zwherea1=zinit)
    return dadt(cosmo, a, zwherea1=zinit)
