"""world_size-P gloo worker (CPU): the slab-decomposed PM protocol (oracle/pm_slab.py -- the same
ghost-plane / transpose / migration scheme as noname_b200/csrc/nnpm_comm.cuh) equals the
single-process oracle.  Launched by tests/test_slab_gloo.py under torch.distributed.run."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    import torch.distributed as dist
    from conftest import make_particles, nlinf
    from oracle import pm_ref as O
    from oracle.pm_slab import SlabPM

    dist.init_process_group("gloo")
    rank, P = dist.get_rank(), dist.get_world_size()
    N = 16
    dims = (N, N, N)
    npart = 6000
    x, v = make_particles(77, npart, 3, clustered=True)
    v *= 0.2
    m = float(N ** 3) / npart
    errs = []

    def check(name, ok, detail=""):
        if not ok:
            errs.append(f"rank {rank}: {name} FAILED {detail}")

    # single-process oracle, 3 comoving + 2 static steps
    f = O.Fields(dims, npart, 3)
    xr, vr = x.copy(), v.copy()
    sc = (1.01, 1.02, 0.8, 1.03)
    dts = (0.02, 0.03, 0.025)
    ref_rho, ref_phi = [], []
    for dt in dts:
        O.step_comoving(f, xr, vr, m, dt, *sc)
        ref_rho.append(f.rho.copy())
        ref_phi.append(f.phi.copy())
    for dt in dts[:2]:
        O.step_static(f, xr, vr, m, dt)
        ref_rho.append(f.rho.copy())
        ref_phi.append(f.phi.copy())

    # slab run: every rank starts with a contiguous chunk of the (unsorted) list, then routes it
    S = SlabPM(dims)
    lo, hi = rank * npart // P, (rank + 1) * npart // P
    xs, vs, ids = S.migrate(x[lo:hi].copy(), v[lo:hi].copy(), np.arange(lo, hi))
    check("ownership", bool(np.all(S.owner(xs) == rank)))
    step = 0
    nmig = 0
    for dt in dts:
        n0 = set(ids.tolist())
        xs, vs, ids = S.step_comoving(xs, vs, ids, m, dt, *sc)
        nmig += len(set(ids.tolist()) - n0)
        sl = slice(S.kz0, S.kz0 + S.nzl)
        # mesh holds phi after the step; rho parity checked through a separate deposit below
        check(f"phi step {step}", np.max(np.abs(S.mesh[2:2 + S.nzl] - ref_phi[step][sl])) < 1e-12 * np.max(np.abs(ref_phi[step])))
        step += 1
    for dt in dts[:2]:
        xs, vs, ids = S.step_static(xs, vs, ids, m, dt)
        sl = slice(S.kz0, S.kz0 + S.nzl)
        check(f"phi step {step}", np.max(np.abs(S.mesh[2:2 + S.nzl] - ref_phi[step][sl])) < 1e-12 * np.max(np.abs(ref_phi[step])))
        step += 1
    check("particles", np.max(np.abs(xs - xr[ids])) < 1e-13 and nlinf(vs, vr[ids]) < 1e-11,
          f"{np.max(np.abs(xs - xr[ids])):.2e} {nlinf(vs, vr[ids]):.2e}")
    xw = xs.copy()
    O.make_periodic(xw)
    check("ownership after steps", bool(np.all(S.owner(xw) == rank)))
    import torch
    cnt = torch.tensor([len(ids), nmig], dtype=torch.int64)
    dist.all_reduce(cnt)
    check("particle count conserved", int(cnt[0]) == npart, str(cnt))
    check("some particles migrated", int(cnt[1]) > 0)
    # deposit parity of the final state (ghost-plane fold)
    rho = S.deposit(xw, m).copy()
    ref = np.zeros((N, N, N))
    xrw = xr.copy()
    O.make_periodic(xrw)
    O.deposit(ref, xrw, m, interpolation="cic")
    check("rho fold", np.max(np.abs(rho - ref[S.kz0:S.kz0 + S.nzl])) < 1e-12 * np.max(ref))
    tot = torch.tensor([rho.sum()], dtype=torch.float64)
    dist.all_reduce(tot)
    check("mass", abs(float(tot) - m * npart) < 1e-9 * m * npart)
    gm = S.global_max(float(np.max(np.abs(vs))))
    check("global max", gm == float(np.max(np.abs(vr))) or abs(gm - np.max(np.abs(vr))) < 1e-12)

    allerrs = [None] * P
    dist.all_gather_object(allerrs, errs)
    flat = [e for es in allerrs for e in es]
    if rank == 0:
        print("\n".join(flat) if flat else "SLAB_GLOO_OK")
    dist.destroy_process_group()
    return 1 if flat else 0


if __name__ == "__main__":
    sys.exit(main())
