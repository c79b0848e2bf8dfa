"""Multi-GPU parity worker: run under torchrun (one process per GPU).  Every rank builds the same
seeded inputs, uploads its share, and the slab-decomposed result is compared with the CPU oracle
(and, for the fixed-point deposit, bit for bit with a single-GPU context on rank 0's device)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    import torch
    import torch.distributed as dist
    from conftest import make_particles, nlinf
    from noname_b200 import DevicePM, NNPMError
    from oracle import pm_ref as O

    rank = int(os.environ["RANK"])
    P = int(os.environ["WORLD_SIZE"])
    lr = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(lr)
    dist.init_process_group("gloo")
    N = 64                     # >= 64: the TMA column FFT, the fused z pass and the fused NVLink transposes all run
    dims = (N, N, N)
    npart = 60000
    x, v = make_particles(101, npart, 3, clustered=True)
    v *= 0.2
    m = float(N ** 3) / npart
    nzl = N // P
    lo, hi = rank * npart // P, (rank + 1) * npart // P
    errs = []

    def check(name, ok, detail=""):
        if not ok:
            errs.append(f"rank {rank}: {name} FAILED {detail}")

    # (deposit mode, sort cadence): sort_every > 0 runs the tiled shared-memory deposit / push on slabs
    # the last case runs the slab transposes on the copy engines alone instead of the default (copy engines + SM kernel)
    for mode, sort_every, opts in (("atomic", 0, {}), ("fixedpoint", 0, {}), ("atomic", 1, {}), ("fixedpoint", 2, {}),
                                   ("atomic", 1, {"peer_transpose": 1})):
        dev = DevicePM(dims, npart, rank=3, device=lr, nranks=P, rank_id=rank, deposit_mode=mode, timing=True,
                       sort_every=sort_every, part_capacity=npart)   # clustered particles: a slab may hold far more than npart / P
        for k_, v_ in opts.items():
            dev.set_option(k_, v_)
        uid = [DevicePM.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        dev.comm_init(uid[0])
        # deliberately upload particles to the "wrong" rank: contiguous chunks of the random list
        dev.upload(np.ascontiguousarray(x[lo:hi]), np.ascontiguousarray(v[lo:hi]), m, first_id=lo)
        dev.redistribute()
        cnt = torch.tensor([dev.local_count], dtype=torch.int64)
        dist.all_reduce(cnt)
        check("redistribute count", int(cnt.item()) == npart, str(cnt))
        xs, vs, ids = dev.download(ids=True)
        k = np.mod(np.floor(np.where(xs[:, 2] > 1, xs[:, 2] - 1, xs[:, 2]) * N).astype(np.int64), N)
        check("ownership", bool(np.all(k // nzl == rank)))
        check("payload", np.array_equal(xs, x[ids]) and np.array_equal(vs, v[ids]))

        # deposit parity (slab)
        dev.deposit()
        rho = dev.get_field("rho")
        ref = np.zeros((N, N, N))
        xw = x.copy()
        O.make_periodic(xw)
        O.deposit(ref, xw, m, "cic")
        check(f"rho {mode}", nlinf(rho, ref[rank * nzl:(rank + 1) * nzl]) < (1e-12 if mode == "atomic" else 1e-11),
              str(nlinf(rho, ref[rank * nzl:(rank + 1) * nzl])))
        if mode == "fixedpoint":
            one = None
            if rank == 0:
                with DevicePM(dims, npart, rank=3, device=lr, deposit_mode=mode) as d1:
                    d1.upload(x, v, m)
                    d1.deposit()
                    one = d1.get_field("rho")
            box = [one]
            dist.broadcast_object_list(box, src=0)
            check("rho fixedpoint bit-exact vs 1 GPU", np.array_equal(rho, box[0][rank * nzl:(rank + 1) * nzl]))

        # three comoving steps vs oracle
        f = O.Fields(dims, npart, 3)
        xr, vr = x.copy(), v.copy()
        dt = 0.02
        for it in range(3):
            sc = (1.0 + 0.01 * it, 1.02 + 0.01 * it, 0.9, 1.03 + 0.01 * it)
            O.step_comoving(f, xr, vr, m, dt, *sc)
            st = dev.step_comoving(dt, *sc)
            dev.reduce_stats(st)
            check(f"max_abs_v step {it}", abs(st.max_abs_v - np.max(np.abs(vr))) <= 1e-11 * np.max(np.abs(vr)))
            check(f"halo {it}", st.n_halo_violations == 0)
        phi = dev.get_field("phi")
        check(f"phi {mode}", nlinf(phi, f.phi[rank * nzl:(rank + 1) * nzl]) < 1e-11, str(nlinf(phi, f.phi[rank * nzl:(rank + 1) * nzl])))
        xs, vs, ids = dev.download(ids=True)
        cnt = torch.tensor([dev.local_count, st.n_migrated], dtype=torch.int64)
        dist.all_reduce(cnt)
        check("count after steps", int(cnt[0].item()) == npart)
        check(f"x {mode}", np.max(np.abs(xs - xr[ids])) < 1e-10, str(np.max(np.abs(xs - xr[ids]))))
        check(f"v {mode}", nlinf(vs, vr[ids]) < 1e-10, str(nlinf(vs, vr[ids])))
        kr, pr = O.powerSpectrum(xr.copy(), m, dims, "cic")
        kg, pg = dev.power_spectrum()
        ok = np.isfinite(pr)
        check("P(k)", np.max(np.abs(pg[ok] / pr[ok] - 1)) < 1e-6)
        if rank == 0:
            print(f"[mgpu P={P} {mode} sort_every={sort_every}] untiled={st.n_untiled} migrated(last step, all ranks)={int(cnt[1].item())} ms={st.as_dict()['ms']}")
        dev.close()

    # the transpose paths -- fused into the FFT kernels' stores, copy engines pipelined with chunked transforms
    # (default), copy engines un-pipelined, the SM bulk-copy transport (alone and pipelined), grouped NCCL send/recv
    # -- equal the oracle and each other bit for bit
    # (N1 = 128 so that the own row-FFT x passes run too)
    N2 = 64
    rng = np.random.default_rng(5)
    rho = rng.random((N2, N2, 2 * N2))
    ref = np.empty_like(rho)
    O.compute_potential(ref, rho - rho.mean(), np.zeros((N2, N2, N2 + 1), dtype=np.complex128))
    nz2 = N2 // P
    got = {}
    for name, opts in (("fused", {"fuse_transpose": 1}), ("default", {}), ("pipelined", {"fuse_transpose": 0, "pipeline": 2}),
                       ("unpipelined", {"fuse_transpose": 0, "pipeline": 0}), ("sm", {"peer_transpose": 2, "pipeline": 0}),
                       ("sm_pipelined", {"peer_transpose": 2, "pipeline": 2}), ("sm_few_ctas", {"peer_transpose": 2, "xfer_ctas": 3}),
                       ("sm_last_chunk", {"peer_transpose": 3, "pipeline": 2}),
                       ("nccl", {"fuse_transpose": 0, "peer_transpose": 0})):
        dev = DevicePM((2 * N2, N2, N2), 8, rank=3, device=lr, nranks=P, rank_id=rank)
        uid = [DevicePM.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        dev.comm_init(uid[0])
        try:
            for k_, v_ in opts.items():
                dev.set_option(k_, v_)
        except NNPMError as ex:   # "fuse_transpose" = 1 exists only in `make EXPERIMENTS=1` builds (same library on every rank)
            assert name == "fused" and "NNPM_EXPERIMENTS" in str(ex), str(ex)
            dev.close()
            continue
        for rep in range(2):
            dev.set_field("rho", np.ascontiguousarray(rho[rank * nz2:(rank + 1) * nz2]))
            dev.solve()
            got[name] = dev.get_field("phi")
        check(f"solve {name}", nlinf(got[name], ref[rank * nz2:(rank + 1) * nz2]) < 1e-12 * np.max(np.abs(ref)) / np.max(np.abs(ref[rank * nz2:(rank + 1) * nz2])))
        dev.close()
    base = "fused" if "fused" in got else "default"
    for name in ("default", "pipelined", "unpipelined", "sm", "sm_pipelined", "sm_few_ctas", "sm_last_chunk", "nccl"):
        check(f"solve {base} == {name}", np.array_equal(got[base], got[name]))

    # fine mesh along z (N3 = 2048, BASELINE config 5): slab transposes + the radix-2-wrapped fused z pass
    d5 = (16, 16, 2048)
    rng = np.random.default_rng(6)
    rho = rng.random((d5[2], d5[1], d5[0]))
    ref = np.empty_like(rho)
    O.compute_potential(ref, rho - rho.mean(), np.zeros((d5[2], d5[1], d5[0] // 2 + 1), dtype=np.complex128))
    nz5 = d5[2] // P
    dev = DevicePM(d5, 8, rank=3, device=lr, nranks=P, rank_id=rank)
    uid = [DevicePM.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(uid, src=0)
    dev.comm_init(uid[0])
    dev.set_field("rho", np.ascontiguousarray(rho[rank * nz5:(rank + 1) * nz5]))
    dev.solve()
    phi5 = dev.get_field("phi")
    check("solve N3=2048", nlinf(phi5, ref[rank * nz5:(rank + 1) * nz5]) < 1e-11, str(nlinf(phi5, ref[rank * nz5:(rank + 1) * nz5])))
    dev.close()

    # device-generated ICs agree across decompositions
    dev = DevicePM(dims, N ** 3, rank=3, device=lr, nranks=P, rank_id=rank)
    uid = [DevicePM.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(uid, src=0)
    dev.comm_init(uid[0])
    dev.init_synthetic(dims, seed=3, nmodes=8, kmax=3, disp_rms=0.05, vfac=0.3, m=1.0)
    xs, vs, ids = dev.download(ids=True)
    xr, vr = O.synthetic_zeldovich(dims, 3, 3, 8, 3, 0.05, 0.3)
    check("synthetic ICs", np.max(np.abs(xs - xr[ids])) < 1e-13 and np.max(np.abs(vs - vr[ids])) < 1e-13)
    dev.close()

    all_errs = [None] * P
    dist.all_gather_object(all_errs, errs)
    flat = [e for l in all_errs for e in l]
    if rank == 0:
        print("MGPU_OK" if not flat else "MGPU_FAIL\n" + "\n".join(flat))
    dist.destroy_process_group()
    return 1 if flat else 0


if __name__ == "__main__":
    sys.exit(main())
