"""Device-resident particle-mesh state: a thin object wrapper over the C ABI (include/nnpm.h).

Array conventions are the reference's Julia column-major memory viewed by numpy in C order with
the axes reversed (see include/nnpm.h):
    rho/phi  (N3,N2,N1)      acc (N3,N2,N1,3)      x/v/pa (Npart, rank)
With nranks > 1 mesh arrays hold the local slab only: (N3/P, N2, N1[,3]).
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from ._lib import NNPMError, Config, Stats


def _f64(a, name):
    if not isinstance(a, np.ndarray) or a.dtype != np.float64 or not a.flags["C_CONTIGUOUS"]:
        raise TypeError(f"{name} must be a C-contiguous float64 numpy array (the reference passes "
                        f"Array{{Float64}}; got {type(a).__name__} {getattr(a, 'dtype', None)})")
    return a


class DevicePM:
    """One context = one GPU = one slab of the mesh (nranks > 1) or all of it."""

    def __init__(self, dims, npart_total, rank=None, deposit="cic", backinterp="cic",
                 deposit_mode="atomic", fixed_bits=0, sort_every=0, bugcompat_interp_z=False,
                 device=0, nranks=1, rank_id=0, timing=False, part_capacity=0, stream=None):
        L = _lib.load()
        dims = tuple(int(d) for d in dims)
        if len(dims) != 3:
            raise ValueError("dims must be (N1, N2, N3) -- TopgridDimensions")
        if rank is None:
            rank = sum(1 for d in dims if d > 1)        # ParticleMesh.jl:105
        for nm, val, tab in (("deposit", deposit, _lib.INTERP), ("backinterp", backinterp, _lib.INTERP),
                             ("deposit_mode", deposit_mode, _lib.DEPOSIT_MODE)):
            if val not in tab:
                raise ValueError(f"{nm} must be one of {sorted(tab)} (got {val!r})")
        self.dims, self.rank, self.npart_total = dims, int(rank), int(npart_total)
        self.nranks, self.rank_id = int(nranks), int(rank_id)
        self.nzl = dims[2] // self.nranks
        cfg = Config()
        cfg.struct_size = C.sizeof(Config)
        cfg.rank = self.rank
        cfg.ngrid[:] = dims
        cfg.npart_total = self.npart_total
        cfg.part_capacity = int(part_capacity)
        cfg.deposit = _lib.INTERP[deposit]
        cfg.backinterp = _lib.INTERP[backinterp]
        cfg.deposit_mode = _lib.DEPOSIT_MODE[deposit_mode]
        cfg.fixed_bits = int(fixed_bits)
        cfg.sort_every = int(sort_every)
        cfg.bugcompat_interp_z = int(bool(bugcompat_interp_z))
        cfg.device = int(device)
        cfg.nranks = self.nranks
        cfg.rank_id = self.rank_id
        cfg.timing = int(bool(timing))
        cfg.stream = stream
        self._L = L
        self._h = C.c_void_p()
        rc = L.nnpm_create(C.byref(self._h), C.byref(cfg))
        if rc != 0:
            raise NNPMError(rc, (L.nnpm_last_error(None) or b"").decode())
        self.m = 0.0

    # -- plumbing -------------------------------------------------------------------------
    def _ck(self, rc):
        if rc != 0:
            raise NNPMError(rc, (self._L.nnpm_last_error(self._h) or b"").decode())

    def close(self):
        if getattr(self, "_h", None):
            self._L.nnpm_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    @property
    def local_count(self):
        return int(self._L.nnpm_local_count(self._h))

    @property
    def device_bytes(self):
        return int(self._L.nnpm_device_bytes(self._h))

    def set_option(self, name, value):
        self._ck(self._L.nnpm_set_option(self._h, name.encode(), float(value)))

    # -- multi-GPU ------------------------------------------------------------------------
    @staticmethod
    def comm_unique_id():
        L = _lib.load()
        buf = (C.c_uint8 * 128)()
        rc = L.nnpm_comm_unique_id(buf)
        if rc != 0:
            raise NNPMError(rc, "nnpm_comm_unique_id failed (NCCL not loadable?)")
        return bytes(buf)

    def comm_init(self, uid: bytes):
        buf = (C.c_uint8 * 128).from_buffer_copy(uid)
        self._ck(self._L.nnpm_comm_init(self._h, buf))

    def redistribute(self):
        self._ck(self._L.nnpm_redistribute(self._h))

    # -- particles ------------------------------------------------------------------------
    def upload(self, x, v=None, m=1.0, first_id=0):
        x = _f64(x, "x")
        if x.ndim != 2 or x.shape[1] != self.rank:
            raise ValueError(f"x must have shape (Npart, {self.rank})")
        if v is not None:
            v = _f64(v, "v")
            if v.shape != x.shape:
                raise ValueError("v must have the shape of x")
        self.m = float(m)
        self._ck(self._L.nnpm_upload_particles(self._h, x.ctypes.data, v.ctypes.data if v is not None else None,
                                               x.shape[0], int(first_id), self.m))

    def download(self, x=None, v=None, ids=False):
        n = self.local_count
        if x is None:
            x = np.empty((n, self.rank))
        if v is None:
            v = np.empty((n, self.rank))
        _f64(x, "x"), _f64(v, "v")
        if x.shape != (n, self.rank) or v.shape != (n, self.rank):
            raise ValueError(f"x and v must have shape ({n}, {self.rank})")
        idarr = np.empty(n, dtype=np.int64) if ids else None
        self._ck(self._L.nnpm_download_particles(self._h, x.ctypes.data, v.ctypes.data,
                                                 idarr.ctypes.data if ids else None))
        return (x, v, idarr) if ids else (x, v)

    def init_synthetic(self, pdims, kind="zeldovich", seed=12345, nmodes=32, kmax=8, disp_rms=0.0,
                       vfac=0.0, m=1.0):
        pd = (C.c_int64 * 3)(*[int(p) for p in pdims])
        k = {"zeldovich": _lib.IC_ZELDOVICH_MODES, "lattice": _lib.IC_LATTICE, "clustered": _lib.IC_CLUSTERED}[kind]
        self.m = float(m)
        self._ck(self._L.nnpm_init_synthetic(self._h, k, pd, int(seed), int(nmodes), int(kmax),
                                             float(disp_rms), float(vfac), self.m))

    def init_grf(self, pdims, seed=12345, ps_index=-1.0, disp_rms=0.0, vfac=0.0, m=1.0, peaks=None):
        """PowerSpectrumParticles on the device, seeded (ParticleMesh.jl:209-254): scaleFreePS(ps_index, .), or -- with
        peaks = (k_list, lnwidth_list) -- multiplePeaksPS(ps_index, ., k_list, lnwidth_list) (InputPowerSpectra.jl:6-36)."""
        pd = (C.c_int64 * 3)(*[int(p) for p in pdims])
        self.m = float(m)
        if peaks is None:
            self._ck(self._L.nnpm_init_grf(self._h, pd, int(seed), float(ps_index), float(disp_rms), float(vfac), self.m))
            return
        ks, ws = [float(q) for q in peaks[0]], [float(q) for q in peaks[1]]
        if len(ks) != len(ws):
            raise ValueError("peaks = (k_list, lnwidth_list) of equal lengths")
        n = len(ks)
        self._ck(self._L.nnpm_init_grf_peaks(self._h, pd, int(seed), float(ps_index), n, (C.c_double * max(n, 1))(*ks),
                                             (C.c_double * max(n, 1))(*ws), float(disp_rms), float(vfac), self.m))

    def init_pancake(self, pdims, xamp, vamp, m=1.0):
        """ZeldovichPancakeParticles on the device (ParticleMesh.jl:177-198): xamp = 2/5 A a, vamp = 2/5 A adot."""
        pd = (C.c_int64 * 3)(*[int(p) for p in pdims])
        self.m = float(m)
        self._ck(self._L.nnpm_init_pancake(self._h, pd, float(xamp), float(vamp), self.m))

    def sort(self):
        self._ck(self._L.nnpm_sort(self._h))

    # -- operators (un-fused) -------------------------------------------------------------
    def make_periodic(self):
        self._ck(self._L.nnpm_make_periodic(self._h))

    def deposit(self):
        self._ck(self._L.nnpm_deposit(self._h))

    def solve(self):
        self._ck(self._L.nnpm_solve(self._h))

    def accel(self):
        self._ck(self._L.nnpm_accel(self._h))

    def interp(self):
        self._ck(self._L.nnpm_interp(self._h))

    def drift(self, dt):
        self._ck(self._L.nnpm_drift(self._h, float(dt)))

    def kick(self, dt, a=None, dadt=None):
        if a is None:
            self._ck(self._L.nnpm_kick(self._h, float(dt)))
        else:
            self._ck(self._L.nnpm_kick_comoving(self._h, float(dt), float(a), float(dadt)))

    # -- fields ---------------------------------------------------------------------------
    def _field_shape(self, which):
        n1, n2, _ = self.dims
        if which in (_lib.FIELD_RHO, _lib.FIELD_PHI):
            return (self.nzl, n2, n1)
        if which == _lib.FIELD_ACC:
            return (self.nzl, n2, n1, 3)
        return (self.local_count, self.rank)

    def get_field(self, which, out=None):
        which = {"rho": 0, "phi": 1, "acc": 2, "pa": 3}.get(which, which)
        shp = self._field_shape(which)
        if out is None:
            out = np.empty(shp)
        _f64(out, "out")
        if out.shape != shp:
            raise ValueError(f"field array must have shape {shp}, got {out.shape}")
        self._ck(self._L.nnpm_get_field(self._h, which, out.ctypes.data))
        return out

    def set_field(self, which, arr):
        which = {"rho": 0, "phi": 1, "acc": 2, "pa": 3}.get(which, which)
        _f64(arr, "arr")
        shp = self._field_shape(which)
        if arr.shape != shp:
            raise ValueError(f"field array must have shape {shp}, got {arr.shape}")
        self._ck(self._L.nnpm_set_field(self._h, which, arr.ctypes.data))

    # -- fused steps ----------------------------------------------------------------------
    def step_static(self, dt):
        st = Stats()
        self._ck(self._L.nnpm_step_static(self._h, float(dt), C.byref(st)))
        return st

    def step_comoving(self, dt, a_d1, a_k, adot_k, a_d2):
        st = Stats()
        self._ck(self._L.nnpm_step_comoving(self._h, float(dt), float(a_d1), float(a_k), float(adot_k),
                                            float(a_d2), C.byref(st)))
        return st

    def reduce_stats(self, st):
        self._ck(self._L.nnpm_reduce_stats(self._h, C.byref(st)))
        return st

    def power_spectrum(self):
        lenk = int(round(max(self.dims) / 2))
        k = np.empty(lenk)
        p = np.empty(lenk)
        dp = C.POINTER(C.c_double)
        self._ck(self._L.nnpm_power_spectrum(self._h, k.ctypes.data_as(dp), p.ctypes.data_as(dp), lenk))
        return k, p


_pinned = {}


def pinned_empty(shape, dtype=np.float64):
    """numpy array backed by cudaHostAlloc memory (nnpm_host_alloc). Release with pinned_free."""
    L = _lib.load()
    count = int(np.prod(shape))
    n = count * np.dtype(dtype).itemsize
    p = C.c_void_p()
    rc = L.nnpm_host_alloc(C.byref(p), n)
    if rc != 0:
        raise NNPMError(rc, "nnpm_host_alloc failed")
    buf = (C.c_char * max(n, 1)).from_address(p.value)
    arr = np.frombuffer(buf, dtype=dtype, count=count).reshape(shape)
    _pinned[p.value] = buf
    return arr


def pinned_free(arr):
    L = _lib.load()
    addr = arr.ctypes.data
    if _pinned.pop(addr, None) is not None:
        L.nnpm_host_free(C.c_void_p(addr))
