"""ctypes binding of libnnpm.so -- the C ABI declared in include/nnpm.h.

This is the Python twin of the `ccall` shim in julia/NoNameB200.jl: same
entry points, same argument order.  There is NO CPU fallback: if the shared library has not
been built (noname_b200/csrc/Makefile, or `__graft_entry__.build()`), importing a compute
path raises; if there is no CUDA device, `nnpm_create` fails with NNPM_ERR_CUDA.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libnnpm.so")

NNPM_MS_COUNT = 10
MS_NAMES = ("deposit", "solve", "push", "sort", "exchange", "comm", "fft_z", "fft_xy", "fft_x", "fft_y")
INTERP = {"none": 0, "cic": 1}
DEPOSIT_MODE = {"atomic": 0, "fixedpoint": 1, "fixed": 1}
FIELD_RHO, FIELD_PHI, FIELD_ACC, FIELD_PA = 0, 1, 2, 3
IC_LATTICE, IC_ZELDOVICH_MODES, IC_CLUSTERED = 0, 1, 2

# every symbol include/nnpm.h declares (tests/test_abi.py checks the header against this list
# and the built library against both)
SYMBOLS = (
    "nnpm_version", "nnpm_create", "nnpm_destroy", "nnpm_last_error", "nnpm_comm_unique_id",
    "nnpm_comm_init", "nnpm_host_alloc", "nnpm_host_free", "nnpm_upload_particles",
    "nnpm_redistribute", "nnpm_local_count", "nnpm_download_particles", "nnpm_init_synthetic", "nnpm_init_grf",
    "nnpm_init_grf_peaks", "nnpm_init_pancake",
    "nnpm_sort", "nnpm_make_periodic", "nnpm_deposit", "nnpm_solve", "nnpm_accel", "nnpm_interp",
    "nnpm_drift", "nnpm_kick", "nnpm_kick_comoving", "nnpm_get_field", "nnpm_set_field",
    "nnpm_step_static", "nnpm_step_comoving", "nnpm_reduce_stats", "nnpm_device_bytes",
    "nnpm_power_spectrum", "nnpm_set_option",
)


class NNPMError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"nnpm error {code}: {msg}")
        self.code = code


class Config(C.Structure):
    _fields_ = [
        ("struct_size", C.c_int32), ("rank", C.c_int32), ("ngrid", C.c_int64 * 3),
        ("npart_total", C.c_int64), ("part_capacity", C.c_int64), ("deposit", C.c_int32),
        ("backinterp", C.c_int32), ("deposit_mode", C.c_int32), ("fixed_bits", C.c_int32),
        ("sort_every", C.c_int32), ("bugcompat_interp_z", C.c_int32), ("device", C.c_int32),
        ("nranks", C.c_int32), ("rank_id", C.c_int32), ("timing", C.c_int32), ("stream", C.c_void_p),
    ]


class Stats(C.Structure):
    _fields_ = [
        ("max_abs_v", C.c_double), ("max_v", C.c_double), ("max_pa", C.c_double),
        ("n_halo_violations", C.c_int64), ("npart_local", C.c_int64), ("n_migrated", C.c_int64),
        ("n_untiled", C.c_int64),
        ("ms", C.c_float * NNPM_MS_COUNT), ("kernel_launches", C.c_int32), ("pad", C.c_int32),
    ]

    def as_dict(self):
        d = {k: getattr(self, k) for k in ("max_abs_v", "max_v", "max_pa", "n_halo_violations",
                                           "npart_local", "n_migrated", "n_untiled", "kernel_launches")}
        d["ms"] = {n: float(self.ms[i]) for i, n in enumerate(MS_NAMES)}
        return d


_lib = None


def load():
    """Load libnnpm.so (once).  Raises if it is missing -- there is no fallback path."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is not built: run `make -C noname_b200/csrc` (or __graft_entry__.build()). "
            "noname_b200 has no CPU fallback.")
    L = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL if hasattr(C, "RTLD_GLOBAL") else 0)
    vp, dp, i64 = C.c_void_p, C.POINTER(C.c_double), C.c_int64
    sig = {
        "nnpm_version": ([], C.c_int),
        "nnpm_create": ([C.POINTER(vp), C.POINTER(Config)], C.c_int),
        "nnpm_destroy": ([vp], C.c_int),
        "nnpm_last_error": ([vp], C.c_char_p),
        "nnpm_comm_unique_id": ([C.POINTER(C.c_uint8)], C.c_int),
        "nnpm_comm_init": ([vp, C.POINTER(C.c_uint8)], C.c_int),
        "nnpm_host_alloc": ([C.POINTER(vp), C.c_uint64], C.c_int),
        "nnpm_host_free": ([vp], C.c_int),
        "nnpm_upload_particles": ([vp, vp, vp, i64, i64, C.c_double], C.c_int),
        "nnpm_redistribute": ([vp], C.c_int),
        "nnpm_local_count": ([vp], i64),
        "nnpm_download_particles": ([vp, vp, vp, vp], C.c_int),
        "nnpm_init_synthetic": ([vp, C.c_int, C.POINTER(i64), C.c_uint64, C.c_int32, C.c_int32,
                                 C.c_double, C.c_double, C.c_double], C.c_int),
        "nnpm_init_grf": ([vp, C.POINTER(i64), C.c_uint64, C.c_double, C.c_double, C.c_double, C.c_double], C.c_int),
        "nnpm_init_grf_peaks": ([vp, C.POINTER(i64), C.c_uint64, C.c_double, C.c_int32, C.POINTER(C.c_double),
                                 C.POINTER(C.c_double), C.c_double, C.c_double, C.c_double], C.c_int),
        "nnpm_init_pancake": ([vp, C.POINTER(i64), C.c_double, C.c_double, C.c_double], C.c_int),
        "nnpm_sort": ([vp], C.c_int),
        "nnpm_make_periodic": ([vp], C.c_int),
        "nnpm_deposit": ([vp], C.c_int),
        "nnpm_solve": ([vp], C.c_int),
        "nnpm_accel": ([vp], C.c_int),
        "nnpm_interp": ([vp], C.c_int),
        "nnpm_drift": ([vp, C.c_double], C.c_int),
        "nnpm_kick": ([vp, C.c_double], C.c_int),
        "nnpm_kick_comoving": ([vp, C.c_double, C.c_double, C.c_double], C.c_int),
        "nnpm_get_field": ([vp, C.c_int, vp], C.c_int),
        "nnpm_set_field": ([vp, C.c_int, vp], C.c_int),
        "nnpm_step_static": ([vp, C.c_double, C.POINTER(Stats)], C.c_int),
        "nnpm_step_comoving": ([vp, C.c_double, C.c_double, C.c_double, C.c_double, C.c_double,
                                C.POINTER(Stats)], C.c_int),
        "nnpm_reduce_stats": ([vp, C.POINTER(Stats)], C.c_int),
        "nnpm_device_bytes": ([vp], i64),
        "nnpm_power_spectrum": ([vp, dp, dp, C.c_int32], C.c_int),
        "nnpm_set_option": ([vp, C.c_char_p, C.c_double], C.c_int),
    }
    assert set(sig) == set(SYMBOLS)
    for name, (args, res) in sig.items():
        f = getattr(L, name)          # AttributeError if the library lacks a declared symbol
        f.argtypes = args
        f.restype = res
    _lib = L
    return L
