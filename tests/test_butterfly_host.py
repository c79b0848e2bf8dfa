"""The radix-R butterfly networks of the own FFT kernels (noname_b200/csrc/nnpm_butterfly.cuh) are plain double2
arithmetic: built here as host C++ with g++ and checked against the DFT definition -- every radix 2 .. 32, both signs,
DIF (natural in, bit-reversed out), DIT (bit-reversed in, natural out), the DIF -> DIT round trip the fused z pass relies
on, and a sub-transform at a register offset.  No GPU needed; the kernels around them are covered by the -m gpu tests."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def harness(tmp_path_factory):
    exe = str(tmp_path_factory.mktemp("bf") / "butterfly_harness")
    src = os.path.join(ROOT, "tests", "host_cpp", "butterfly_harness.cpp")
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-ffp-contract=on", "-o", exe, src, "-lm"])
    out = subprocess.run([exe], capture_output=True, text=True, check=True).stdout
    return [line.split() for line in out.strip().splitlines()]


def test_twiddle_table(harness):
    (row,) = [r for r in harness if r[2] == "twiddles"]
    assert float(row[3]) < 2e-16


@pytest.mark.parametrize("kind,tol_per_r", [("dif", 4e-16), ("dit", 4e-16), ("roundtrip", 8e-16), ("base", 0.0)])
def test_butterflies_match_the_dft(harness, kind, tol_per_r):
    rows = [r for r in harness if r[2] == kind]
    assert len(rows) == 10                                   # R = 2, 4, 8, 16, 32 x sign -1, +1
    for r, sign, _, err in rows:
        # inputs are O(1), outputs O(sqrt(R)): rounding grows with log2(R) stages on sums of R terms
        assert float(err) <= tol_per_r * int(r) * 2, (r, sign, kind, err)
