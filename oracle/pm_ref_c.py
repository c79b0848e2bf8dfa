"""CPU ORACLE driver (test infrastructure, not product code): the C restatement
`oracle/pm_ref.c` bound through ctypes, with scipy's pocketfft standing in for Julia Base
FFTW between the loops (`rfft`/`irfft`, ParticleMesh.jl:481,502).  Same memory layout as
oracle/pm_ref.py.  PARITY UNPINNED (see oracle/pm_ref.py header).

threads == 1 is the faithful single-threaded reference behaviour (Readme.md:53).  threads > 1
(bench.py --impl reference only) runs disjoint chunks of each loop on a thread pool (ctypes
drops the GIL) and pocketfft with workers=threads; the deposit then uses one private rho per
thread summed afterwards.

Only tests/, __graft_entry__.smoke() and bench.py (cpu_baseline, --impl reference) use it.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from concurrent.futures import ThreadPoolExecutor

import numpy as np
import scipy.fft as sfft

_HERE = os.path.dirname(os.path.abspath(__file__))
_dp = C.POINTER(C.c_double)
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
            "pmref_ngpdensity": [_dp, _dp, C.c_double, _i64, C.c_int, _i64, _i64, _i64],
            "pmref_sub": [_dp, _dp, C.c_double, _i64],
            "pmref_add_into": [_dp, _dp, _i64],
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
        rho[...] = 0.0
        if self.pool is None:
            f(_ptr(rho), _ptr(x), m, x.shape[0], rank, n1, n2, n3)
            return
        ch = _chunks(x.shape[0], self.threads)
        priv = [rho] + [np.zeros_like(rho) for _ in ch[1:]]
        list(self.pool.map(lambda t: f(_ptr(priv[t[0]]), _ptr(x, rank * t[1][0]), m,
                                       t[1][1] - t[1][0], rank, n1, n2, n3), enumerate(ch)))
        for p in priv[1:]:
            self._par(lambda lo, hi, p=p: self.lib.pmref_add_into(_ptr(rho, lo), _ptr(p, lo), hi - lo),
                      rho.size)

    def sub_mean(self, out, rho):
        mean = self.lib.pmref_sum(_ptr(rho), rho.size) / rho.size
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
        a = C.c_double()
        s = C.c_double()
        self.lib.pmref_maxima(_ptr(v), v.size, C.byref(a), C.byref(s))
        return a.value, s.value

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
