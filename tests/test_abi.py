"""The C-ABI boundary (include/nnpm.h <-> noname_b200/libnnpm.so <-> the ctypes binding):
every declared symbol is exported and bound, struct layouts agree, and without a GPU the library
fails loudly instead of falling back to a CPU path.  No compute call is made here."""
import ctypes as C
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "nnpm.h")


def _declared_symbols():
    src = open(HEADER, encoding="utf-8").read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(nnpm_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_match_binding_list():
    from noname_b200 import _lib
    assert _declared_symbols() == sorted(_lib.SYMBOLS)


def test_library_exports_every_declared_symbol(lib_built):
    from noname_b200 import _lib
    out = subprocess.run(["nm", "-D", "--defined-only", _lib.LIB_PATH], capture_output=True, text=True, check=True).stdout
    exported = set(re.findall(r"\sT\s+(nnpm_[a-z0-9_]+)$", out, flags=re.M))
    missing = set(_declared_symbols()) - exported
    assert not missing, f"libnnpm.so lacks {sorted(missing)}"
    for name in _declared_symbols():
        assert hasattr(lib_built, name)


def test_library_is_sm100a_cuda_not_a_cpu_build(lib_built):
    """the shared object carries sm_100a device code (cuobjdump lists the ELF)"""
    from noname_b200 import _lib
    r = subprocess.run(["cuobjdump", "-lelf", _lib.LIB_PATH], capture_output=True, text=True)
    if r.returncode != 0:
        pytest.skip("cuobjdump not available")
    assert "sm_100a" in r.stdout


def test_struct_layouts_match_header():
    """compile a tiny C program against the header and compare sizeof/offsetof with ctypes"""
    from noname_b200 import _lib
    prog = r"""
#include <stdio.h>
#include <stddef.h>
#include "nnpm.h"
int main(void) {
  printf("%zu %zu %zu %zu %zu %zu %zu %zu\n", sizeof(nnpm_config), offsetof(nnpm_config, ngrid),
         offsetof(nnpm_config, part_capacity), offsetof(nnpm_config, stream), sizeof(nnpm_stats),
         offsetof(nnpm_stats, n_untiled), offsetof(nnpm_stats, ms), offsetof(nnpm_stats, kernel_launches));
  return 0;
}
"""
    import tempfile
    with tempfile.TemporaryDirectory() as td:
        src = os.path.join(td, "t.c")
        exe = os.path.join(td, "t")
        open(src, "w").write(prog)
        subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), src, "-o", exe], check=True)
        got = [int(t) for t in subprocess.run([exe], capture_output=True, text=True, check=True).stdout.split()]
    Cf, St = _lib.Config, _lib.Stats
    want = [C.sizeof(Cf), Cf.ngrid.offset, Cf.part_capacity.offset, Cf.stream.offset, C.sizeof(St),
            St.n_untiled.offset, St.ms.offset, St.kernel_launches.offset]
    assert got == want


def test_header_is_plain_c_and_cites_the_reference():
    src = open(HEADER, encoding="utf-8").read()
    code = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    assert "torch" not in code.lower() and "std::" not in code and "template" not in code
    assert 'extern "C"' in src
    for cite in ("ParticleMesh.jl:289", "ParticleMesh.jl:417", "ParticleMesh.jl:631", "ParticleMesh.jl:640",
                 "ParticleMesh.jl:723", "ParticleMesh.jl:800", "ParticleMesh.jl:766", "ParticleMesh.jl:853"):
        assert cite in src, cite


def test_no_cpu_fallback_create_fails_loudly_without_gpu(lib_built):
    """nnpm_create must return NNPM_ERR_CUDA (not silently run on the host) when no device exists."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from noname_b200 import DevicePM, NNPMError
    with pytest.raises(NNPMError) as ei:
        DevicePM((8, 8, 8), 8, rank=3)
    assert ei.value.code == -2 and "no CPU fallback" in str(ei.value)
    from noname_b200 import ParticleMesh as PM
    import numpy as np
    with pytest.raises(NNPMError):
        PM.deposit(np.zeros((4, 4, 4)), np.zeros((3, 3)), 1.0, interpolation="cic")


def test_abi_mismatch_and_bad_arguments_are_rejected(lib_built):
    from noname_b200 import _lib
    cfg = _lib.Config()
    cfg.struct_size = 4
    h = C.c_void_p()
    assert lib_built.nnpm_create(C.byref(h), C.byref(cfg)) == -1
    assert b"ABI mismatch" in lib_built.nnpm_last_error(None)
    cfg.struct_size = C.sizeof(_lib.Config)
    cfg.rank = 4
    assert lib_built.nnpm_create(C.byref(h), C.byref(cfg)) == -1
    assert lib_built.nnpm_version() == 2
    assert lib_built.nnpm_destroy(None) == 0


def test_product_package_never_imports_the_oracle():
    """only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may touch oracle/"""
    pkg = os.path.join(ROOT, "noname_b200")
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dp, f), encoding="utf-8").read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", txt, flags=re.M), f
                assert "pmref_" not in txt and "libpmref" not in txt, f


def test_every_runtime_option_is_documented_in_the_header():
    """nnpm_set_option's accepted names (the strcmp chain in nnpm.cu) == the names the header documents"""
    cu = open(os.path.join(ROOT, "noname_b200", "csrc", "nnpm.cu"), encoding="utf-8").read()
    body = cu[cu.index('extern "C" int nnpm_set_option'):]
    accepted = set(re.findall(r'!strcmp\(name, "([a-z0-9_]+)"\)', body))
    hdr = open(HEADER, encoding="utf-8").read()
    doc = hdr[hdr.index("Runtime tuning knobs"):hdr.index("int nnpm_set_option")]
    documented = set(re.findall(r'"([a-z0-9_]+)"', doc))
    assert accepted, "no options parsed from nnpm.cu"
    assert accepted == documented, f"undocumented: {sorted(accepted - documented)}; stale: {sorted(documented - accepted)}"


HOT_KERNELS = ("k_push_tile", "k_deposit_tile", "k_zfused_s", "k_zfused_p", "k_colfft", "k_rowfft", "k_sort_count", "k_permute")


def test_hot_kernels_do_not_spill(lib_built):
    """ptxas -v output of the in-tree build: no hot kernel spills registers to local memory"""
    log = os.path.join(ROOT, "noname_b200", "csrc", "build.log")
    if not os.path.exists(log):
        pytest.skip("no build.log (library built elsewhere)")
    txt = open(log, encoding="utf-8", errors="replace").read()
    entries = re.findall(r"Compiling entry function '(\w+)'.*?\n.*?\n\s+(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads", txt)
    assert entries, "no ptxas -v records in build.log"
    seen = set()
    for mangled, stack, st, ld in entries:
        hot = [k for k in HOT_KERNELS if k in mangled]
        if not hot:
            continue
        seen.add(hot[0])
        assert int(st) == 0 and int(ld) == 0, f"{mangled} spills ({st} B stores, {ld} B loads)"
    assert seen == set(HOT_KERNELS)


def test_sass_carries_the_blackwell_data_movement_instructions(lib_built):
    """the hot kernels really use the TMA / bulk-copy engine and mbarriers: UTMALDG (tensor loads of the phi box and
    of the FFT column groups), UTMASTG (tensor stores of the column FFT), UBLKRED (bulk async reduction flushing the
    deposit box), UBLKPF (bulk L2 prefetch of the particle runs), SYNCS (mbarrier arrive / try_wait)"""
    from noname_b200 import _lib
    r = subprocess.run(["cuobjdump", "-sass", _lib.LIB_PATH], capture_output=True, text=True)
    if r.returncode != 0:
        pytest.skip("cuobjdump not available")
    per_kernel = {}
    cur = None
    for line in r.stdout.splitlines():
        m = re.search(r"Function : (\w+)", line)
        if m:
            cur = m.group(1)
            continue
        m = re.search(r"\b(UTMALDG|UTMASTG|UBLKRED|UBLKPF|UBLKCP|SYNCS)\b", line)
        if m and cur:
            per_kernel.setdefault(cur, set()).add(m.group(1))

    def ops(name):
        out = set()
        for k, v in per_kernel.items():
            if name in k:
                out |= v
        return out

    assert {"UTMALDG", "SYNCS", "UBLKPF"} <= ops("k_push_tile")
    assert {"UBLKRED", "UBLKPF"} <= ops("k_deposit_tile")
    assert {"UTMALDG", "SYNCS"} <= ops("k_zfused_p")
    assert {"UTMALDG", "SYNCS"} <= ops("k_zfused_s")
    assert {"UTMALDG", "UTMASTG", "SYNCS"} <= ops("k_colfft")
    assert {"UBLKCP", "SYNCS"} <= ops("k_rowfft")
