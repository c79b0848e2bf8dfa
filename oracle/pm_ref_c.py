"""CPU ORACLE driver (test infrastructure, not product code): the C restatement
`oracle/pm_ref.c` bound through ctypes, with scipy's pocketfft standing in for Julia Base
FFTW between the loops (`rfft`/`irfft`, ParticleMesh.jl:481,502).  Same memory layout as
oracle/pm_ref.py.  PARITY UNPINNED (see oracle/pm_ref.py header).

threads == 1 is the faithful single-threaded reference behaviour (Readme.md:53).  threads > 1
(bench.py --impl reference, and the 512^3 field test's checker) runs disjoint chunks of each loop on a
thread pool (ctypes drops the GIL) and pocketfft with workers=threads; the 3-D CIC deposit bins the particles by
z-slab and lets every thread scatter into its own planes (no atomics, no private copies of rho), the sum behind
the mean and the maxima are reduced from per-chunk partials.  Nothing in the threaded path is serial except
Python's dispatch and the prefix sum over threads x threads counters.

Only tests/, __graft_entry__.smoke() and bench.py (cpu_baseline, --impl reference) use it.
"""
from __future__ import annotations

import ctypes as C
import math
import os
import subprocess
from concurrent.futures import ThreadPoolExecutor

import numpy as np
import scipy.fft as sfft

_HERE = os.path.dirname(os.path.abspath(__file__))
_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int64)
_i64 = C.c_int64


def build():
    subprocess.check_call(["make", "-s", "-C", _HERE, "all"])


def _ptr(a, off=0):
    assert a.dtype == np.float64 or a.dtype == np.complex128
    assert a.flags["C_CONTIGUOUS"]
    return C.cast(a.ctypes.data + 8 * off, _dp)


def _chunks(n, parts):
    parts = max(1, min(parts, n))
    b = [n * i // parts for i in range(parts + 1)]
    return [(b[i], b[i + 1]) for i in range(parts) if b[i + 1] > b[i]]


class PMRefC:
    def __init__(self, threads=1):
        path = os.path.join(_HERE, "libpmref.so")
        if not os.path.exists(path):
            build()
        self.lib = L = C.CDLL(path)
        self.threads = int(threads)
        self.pool = ThreadPoolExecutor(self.threads) if self.threads > 1 else None
        sig = {
            "pmref_make_periodic": [_dp, _i64],
            "pmref_cicdensity": [_dp, _dp, C.c_double, _i64, C.c_int, _i64, _i64, _i64],
            "pmref_cicdensity_atomic": [_dp, _dp, C.c_double, _i64, C.c_int, _i64, _i64, _i64],
            "pmref_ngpdensity": [_dp, _dp, C.c_double, _i64, C.c_int, _i64, _i64, _i64],
            "pmref_slab_table": [_i64, _i64, C.c_void_p],
            "pmref_slab_count": [_dp, _i64, _i64, _i64, _i64, C.c_void_p, _ip],
            "pmref_slab_fill": [_dp, _i64, _i64, _i64, _i64, C.c_void_p, _ip, _ip],
            "pmref_cicdensity_slab": [_dp, _dp, _dp, _ip, _i64, C.c_double, _i64, _i64, _i64, _i64],
            "pmref_sub": [_dp, _dp, C.c_double, _i64],
            "pmref_add_into": [_dp, _dp, _i64],
            "pmref_zero": [_dp, _i64],
            "pmref_green": [_dp, _i64, _i64, _i64, _i64, _i64],
            "pmref_scale": [_dp, _i64, C.c_double],
            "pmref_compute_acceleration": [_dp, _dp, _i64, _i64, _i64, _i64, _i64],
            "pmref_cic_interp": [_dp, _dp, _dp, _i64, C.c_int, _i64, _i64, _i64],
            "pmref_ngp_interp": [_dp, _dp, _dp, _i64, C.c_int, _i64, _i64, _i64],
            "pmref_drift": [_dp, _dp, C.c_double, _i64],
            "pmref_kick": [_dp, _dp, C.c_double, _i64],
            "pmref_kick_comoving": [_dp, _dp, C.c_double, C.c_double, C.c_double, _i64],
            "pmref_maxima": [_dp, _i64, _dp, _dp],
        }
        for name, args in sig.items():
            f = getattr(L, name)
            f.argtypes = args
            f.restype = None
        L.pmref_sum.argtypes = [_dp, _i64]
        L.pmref_sum.restype = C.c_double

    def _par(self, fn, n):
        """Run fn(lo, hi) over [0,n): serially, or chunked on the pool."""
        if self.pool is None:
            fn(0, n)
        else:
            list(self.pool.map(lambda c: fn(*c), _chunks(n, self.threads)))

    # -- operator mirrors (same signatures as oracle/pm_ref.py) --------------------------
    def make_periodic(self, x):
        self._par(lambda lo, hi: self.lib.pmref_make_periodic(_ptr(x, lo), hi - lo), x.size)

    def deposit(self, rho, x, m, interpolation="none"):
        n3, n2, n1 = rho.shape
        rank = x.shape[1]
        f = self.lib.pmref_cicdensity if interpolation == "cic" else self.lib.pmref_ngpdensity
        if self.pool is None:
            rho[...] = 0.0
            f(_ptr(rho), _ptr(x), m, x.shape[0], rank, n1, n2, n3)
            return
        self._par(lambda lo, hi: self.lib.pmref_zero(_ptr(rho, lo), hi - lo), rho.size)                # :290-292
        ch = _chunks(x.shape[0], self.threads)
        if not ch or x.shape[0] == 0:
            return
        if interpolation == "cic" and rank == 3 and n3 >= 2 * self.threads:
            self._deposit_slabs(rho, x, m, ch)
            return
        if interpolation == "cic":
            g = self.lib.pmref_cicdensity_atomic
            list(self.pool.map(lambda c: g(_ptr(rho), _ptr(x, rank * c[0]), m, c[1] - c[0], rank, n1, n2, n3), ch))
            return
        priv = [rho] + [np.zeros_like(rho) for _ in ch[1:]]
        list(self.pool.map(lambda t: f(_ptr(priv[t[0]]), _ptr(x, rank * t[1][0]), m,
                                       t[1][1] - t[1][0], rank, n1, n2, n3), enumerate(ch)))
        for p in priv[1:]:
            self._par(lambda lo, hi, p=p: self.lib.pmref_add_into(_ptr(rho, lo), _ptr(p, lo), hi - lo),
                      rho.size)

    def _deposit_slabs(self, rho, x, m, ch):
        """threads > 1, 3-D CIC: stable binning of the particle indices by owning z-slab, then one scatter per slab"""
        n3, n2, n1 = rho.shape
        T, L = self.threads, self.lib
        npart = x.shape[0]

        def ip(a, off=0):
            return C.cast(a.ctypes.data + 8 * off, _ip)
        owner = np.empty(n3, dtype=np.int32)
        L.pmref_slab_table(n3, T, owner.ctypes.data)
        counts = np.zeros((len(ch), T), dtype=np.int64)                  # [chunk][slab]
        list(self.pool.map(lambda t: L.pmref_slab_count(_ptr(x), t[1][0], t[1][1], n3, T, owner.ctypes.data, ip(counts, T * t[0])),
                           enumerate(ch)))
        # idx is ordered by (slab, chunk, position in chunk) = (slab, original index): stable
        flat = counts.T.reshape(-1)                                      # slab-major
        off = np.concatenate(([0], np.cumsum(flat)[:-1])).reshape(T, len(ch)).T.copy()
        idx = np.empty(npart, dtype=np.int64)
        list(self.pool.map(lambda t: L.pmref_slab_fill(_ptr(x), t[1][0], t[1][1], n3, T, owner.ctypes.data, ip(off, T * t[0]), ip(idx)),
                           enumerate(ch)))
        per = counts.sum(axis=0)
        first = np.concatenate(([0], np.cumsum(per)[:-1]))
        ghost = np.zeros((T, n2, n1))
        bounds = [(n3 * t // T, n3 * (t + 1) // T) for t in range(T)]
        list(self.pool.map(lambda t: L.pmref_cicdensity_slab(_ptr(rho), _ptr(ghost, n1 * n2 * t), _ptr(x), ip(idx, int(first[t])),
                                                             int(per[t]), m, n1, n2, n3, bounds[t][1]), range(T)))
        for t in range(T):
            rho[bounds[t][1] % n3] += ghost[t]

    def sub_mean(self, out, rho):
        if self.pool is None:
            mean = self.lib.pmref_sum(_ptr(rho), rho.size) / rho.size
        else:
            parts = list(self.pool.map(lambda c: self.lib.pmref_sum(_ptr(rho, c[0]), c[1] - c[0]), _chunks(rho.size, self.threads)))
            mean = math.fsum(parts) / rho.size
        self._par(lambda lo, hi: self.lib.pmref_sub(_ptr(out, lo), _ptr(rho, lo), mean, hi - lo), rho.size)
        return mean

    def compute_potential(self, phi, rho, rt):
        n3, n2, n1 = rho.shape
        w = self.threads
        rt[...] = sfft.rfftn(rho, workers=w)
        self._par(lambda lo, hi: self.lib.pmref_green(_ptr(rt), n1, n2, n3, lo, hi), n3)
        phi[...] = sfft.irfftn(rt, s=rho.shape, workers=w)
        self._par(lambda lo, hi: self.lib.pmref_scale(_ptr(phi, lo), hi - lo, float(n1)), phi.size)

    def compute_acceleration(self, a, phi):
        n3, n2, n1 = phi.shape
        self._par(lambda lo, hi: self.lib.pmref_compute_acceleration(_ptr(a), _ptr(phi), n1, n2, n3, lo, hi), n3)

    def interpVecToPoints(self, pa, a, x, interpolation="none"):
        n3, n2, n1, _ = a.shape
        rank = x.shape[1]
        f = self.lib.pmref_cic_interp if interpolation == "cic" else self.lib.pmref_ngp_interp
        self._par(lambda lo, hi: f(_ptr(pa, rank * lo), _ptr(a), _ptr(x, rank * lo), hi - lo, rank,
                                   n1, n2, n3), x.shape[0])

    def drift(self, x, v, dt):
        self._par(lambda lo, hi: self.lib.pmref_drift(_ptr(x, lo), _ptr(v, lo), dt, hi - lo), x.size)

    def kick(self, v, a, dt, ascale=None, dadt=None):
        if ascale is None:
            self._par(lambda lo, hi: self.lib.pmref_kick(_ptr(v, lo), _ptr(a, lo), dt, hi - lo), v.size)
        else:
            self._par(lambda lo, hi: self.lib.pmref_kick_comoving(_ptr(v, lo), _ptr(a, lo), dt, ascale,
                                                                  dadt, hi - lo), v.size)

    def maxima(self, v):
        def one(c):
            a, s = C.c_double(), C.c_double()
            self.lib.pmref_maxima(_ptr(v, c[0]), c[1] - c[0], C.byref(a), C.byref(s))
            return a.value, s.value
        if self.pool is None:
            return one((0, v.size))
        parts = list(self.pool.map(one, _chunks(v.size, self.threads)))
        return max(p[0] for p in parts), max(p[1] for p in parts)

    # -- step bodies ----------------------------------------------------------------------
    def step_static(self, f, x, v, m, dt, dep="cic", back="cic"):
        """ParticleMesh.jl:766-778 (+ the max reductions of :782)"""
        self.make_periodic(x)
        self.deposit(f.rho, x, m, dep)
        tmp = np.empty_like(f.rho)
        self.sub_mean(tmp, f.rho)
        self.compute_potential(f.phi, tmp, f.rt)
        self.compute_acceleration(f.acc, f.phi)
        self.interpVecToPoints(f.pa, f.acc, x, back)
        self.drift(x, v, dt / 2)
        self.kick(v, f.pa, dt)
        self.drift(x, v, dt / 2)
        return self.maxima(v)

    def step_comoving(self, f, x, v, m, dt, a_d1, a_k, adot_k, a_d2, dep="cic", back="cic"):
        """ParticleMesh.jl:853-868 (+ the max reductions of :875-878)"""
        self.drift(x, v, dt / 2 / a_d1)
        self.make_periodic(x)
        self.deposit(f.rho, x, m, dep)
        tmp = np.empty_like(f.rho)
        self.sub_mean(tmp, f.rho)
        self.compute_potential(f.phi, tmp, f.rt)
        self.compute_acceleration(f.acc, f.phi)
        self.interpVecToPoints(f.pa, f.acc, x, back)
        self.kick(v, f.pa, dt, a_k, adot_k)
        self.drift(x, v, dt / 2 / a_d2)
        self.make_periodic(x)
        self.maxima(f.pa)
        return self.maxima(v)
