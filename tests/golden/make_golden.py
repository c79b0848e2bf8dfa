"""Generates the committed golden fixtures in this directory.

    python tests/golden/make_golden.py

The reference (Julia 0.4-era source, no tests, no vectors -- SURVEY.md F1/F4) cannot run in this
image, so these vectors come from the CPU oracle (oracle/pm_ref.py), whose per-function fidelity is
pinned by tests/test_oracle_kats.py, plus closed-form answers computed here independently of the
oracle (`analytic_*` entries).  PARITY UNPINNED with respect to the reference's own outputs.
Inputs are seeded; fixtures are small (float64 .npz, <= 0.5 MB each).
"""
import math
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import pm_ref as O  # noqa: E402


def particles(seed, npart, rank):
    rng = np.random.default_rng(seed)
    x = rng.random((npart, rank))
    nb = npart // 3
    c = rng.random((3, rank))
    x[:nb] = np.mod(c[rng.integers(0, 3, nb)] + 0.02 * rng.standard_normal((nb, rank)), 1.0)
    v = 0.01 * rng.standard_normal((npart, rank))
    x[0] = 0.0
    x[1] = np.nextafter(1.0, 0.0)
    x[2] = 1.0
    x[3] = 0.5
    return x, v


def step_case(name, dims, rank, npart, seed, comoving, nsteps):
    x, v = particles(seed, npart, rank)
    m = float(np.prod(dims)) / npart * 0.3
    f = O.Fields(dims, npart, rank)
    out = {"dims": np.array(dims), "rank": rank, "m": m, "x0": x.copy(), "v0": v.copy(), "comoving": int(comoving)}
    dts = [0.01 * (1 + 0.3 * s) for s in range(nsteps)]
    scal = [(1.0 + 0.01 * s + 0.002, 1.0 + 0.01 * s + 0.004, 0.81 - 0.01 * s, 1.0 + 0.01 * s + 0.006) for s in range(nsteps)]
    out["dts"] = np.array(dts)
    out["scalars"] = np.array(scal)
    for s in range(nsteps):
        if comoving:
            O.step_comoving(f, x, v, m, dts[s], *scal[s])
        else:
            O.step_static(f, x, v, m, dts[s])
        if s == 0:
            out["rho1"], out["phi1"], out["acc1"], out["pa1"] = f.rho.copy(), f.phi.copy(), f.acc.copy(), f.pa.copy()
            out["x1"], out["v1"] = x.copy(), v.copy()
    out["xN"], out["vN"], out["phiN"] = x.copy(), v.copy(), f.phi.copy()
    xw = x.copy()
    k, p = O.powerSpectrum(xw, m, dims)
    out["pk_k"], out["pk_p"] = k, p
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(name, {k_: getattr(v_, "shape", v_) for k_, v_ in out.items() if k_ in ("x0", "rho1", "pk_p")})


def analytic_case():
    """closed forms, independent of the oracle: single Fourier mode -> phi, acc (KAT 3);
    one particle -> the eight CIC weights (KAT 1)."""
    N = 16
    nx, ny, nz = 1, 2, 3
    k, j, i = np.meshgrid(np.arange(N), np.arange(N), np.arange(N), indexing="ij")
    phase = 2 * math.pi * (nx * i + ny * j + nz * k) / N
    rho = np.cos(phase)
    eig = -4.0 * N ** 2 * sum(math.sin(math.pi * q / N) ** 2 for q in (nx, ny, nz))
    phi = rho / eig
    acc = np.stack([(1.0 / eig) * N * math.sin(2 * math.pi * q / N) * np.sin(phase) for q in (nx, ny, nz)], axis=-1)
    xp = np.array([[(3 + 0.125) / N, (15 + 0.5) / N, (7 + 0.875) / N]])
    w = np.zeros((N, N, N))
    for (kk, wz) in ((7, 0.125), (8, 0.875)):
        for (jj, wy) in ((15, 0.5), (0, 0.5)):
            for (ii, wx) in ((3, 0.875), (4, 0.125)):
                w[kk, jj, ii] = 2.5 * wx * wy * wz
    np.savez_compressed(os.path.join(HERE, "analytic_16.npz"), rho=rho, phi=phi, acc=acc, xp=xp, mp=2.5, wp=w)
    print("analytic_16")


IC_CASES = {   # name: (pdims, rank, seed, ps_index, disp_rms, vfac, peaks)
    "grf2d": ((12, 10, 1), 2, 4242, -1.0, 0.02, 0.7, None),
    "grf3d": ((8, 6, 4), 3, 99, -2.0, 0.01, 0.0, None),
    "peaks2d": ((16, 12, 1), 2, 777, -1.0, 0.015, 0.5, ([0.1, 1.0, 10.0], [1.0, 1.0, 1.0])),   # the run/PM spectrum
}


def ic_case():
    """seeded initial conditions: the counter-based draws, the Hermitian planes and the window test of multiplePeaksPS
    are pinned against accidental change of the generators' conventions (device twins: nnpm_init_grf[_peaks], nnpm_init_pancake)"""
    out = {}
    for name, (pd, rank, seed, n, rms, vfac, peaks) in IC_CASES.items():
        x, v = O.synthetic_grf(pd, rank, seed, n, rms, vfac, peaks=peaks)
        out[name + "_x"], out[name + "_v"] = x, v
    x, v = O.zeldovich_pancake((16, 8, 1), 2, 10.0, 400.0, 1.0, math.sqrt(2.0 / 3.0))
    out["pancake_x"], out["pancake_v"] = x, v
    np.savez_compressed(os.path.join(HERE, "ics.npz"), **out)
    print("ics", sorted(out))


if __name__ == "__main__":
    if sys.argv[1:] == ["ics"]:
        ic_case()
        sys.exit(0)
    step_case("step3d_comoving_12x10x16", (12, 10, 16), 3, 1500, 101, True, 4)
    step_case("step3d_static_16", (16, 16, 16), 3, 2000, 102, False, 3)
    step_case("step2d_comoving_24x20", (24, 20, 1), 2, 1200, 103, True, 4)
    step_case("step2d_static_32", (32, 32, 1), 2, 1500, 104, False, 3)
    analytic_case()
    ic_case()
