"""world_size-2 gloo worker (CPU): the host side of a slab-decomposed run -- conf key NGPUs, per-rank upload +
redistribute, reduced step statistics driving dt, gathers back into the reference's gp / gd / p at outputs, rank-0
output files -- exercised through noname_b200.ParticleMesh.evolvePMCosmology with the device context replaced by an
adapter over the CPU slab protocol (oracle/pm_slab.py).  The result must equal the single-process oracle loop.
Launched by tests/test_slab_gloo.py under torch.distributed.run."""
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


class _Stats:
    def __init__(self, mabs, mv, mpa):
        self.max_abs_v, self.max_v, self.max_pa = mabs, mv, mpa


def make_adapter():
    import torch.distributed as dist
    from oracle.pm_slab import SlabPM, GLO
    from noname_b200 import _lib

    class SlabDev:
        """DevicePM's interface (the part the evolve loops use) over oracle.pm_slab.SlabPM"""

        def __init__(self, dims, npart_total, rank=3, nranks=1, rank_id=0, **kw):
            assert nranks == dist.get_world_size() and rank_id == dist.get_rank()
            self.s = SlabPM(tuple(dims))
            self.dims, self.rank, self.nranks, self.rank_id = tuple(dims), rank, nranks, rank_id
            self.keep, self.rho = False, None
            self.inited = False

        def comm_init(self, uid):
            assert isinstance(uid, bytes) and len(uid) == 128
            self.inited = True

        def set_option(self, name, value):
            assert name == "keep_rho"
            self.keep = bool(value)

        def upload(self, x, v, m, first_id=0):
            self.x, self.v, self.m = x.copy(), v.copy(), m
            self.ids = np.arange(first_id, first_id + len(x), dtype=np.int64)

        def redistribute(self):
            assert self.inited
            self.x, self.v, self.ids = self.s.migrate(self.x, self.v, self.ids)

        @property
        def local_count(self):
            return len(self.x)

        def step_comoving(self, dt, a1, ak, adk, a2):
            from oracle import pm_ref as O
            x, v = self.x, self.v
            O.drift(x, v, dt / 2 / a1)
            O.make_periodic(x)
            rho = self.s.deposit(x, self.m)
            self.rho = rho.copy() if self.keep else None
            self.s.solve()
            pa = self.s.gather(x)
            O.kick(v, pa, dt, ak, adk)
            O.drift(x, v, dt / 2 / a2)
            O.make_periodic(x)
            st = _Stats(np.max(np.abs(v)) if len(v) else 0.0, np.max(v) if len(v) else -np.inf, np.max(pa) if len(pa) else -np.inf)
            self.x, self.v, self.ids = self.s.migrate(x, v, self.ids)
            return st

        def reduce_stats(self, st):
            return _Stats(self.s.global_max(st.max_abs_v), self.s.global_max(st.max_v), self.s.global_max(st.max_pa))

        def download(self, x=None, v=None, ids=False):
            return (self.x.copy(), self.v.copy(), self.ids.copy()) if ids else (self.x.copy(), self.v.copy())

        def get_field(self, which):
            if which == "phi":
                return self.s.mesh[GLO:GLO + self.s.nzl].copy()
            if self.rho is None:
                raise _lib.NNPMError(-6, "mesh does not currently hold rho")
            return self.rho.copy()

        def deposit(self):
            raise AssertionError("the loop must have kept the density of the last step")

        @staticmethod
        def comm_unique_id():
            return bytes(128)

        def close(self):
            pass

    return SlabDev


def main():
    import torch.distributed as dist
    from noname_b200 import ParticleMesh as PM, conf as cf
    from oracle import pm_ref as O

    dist.init_process_group("gloo")
    rank, P = dist.get_rank(), dist.get_world_size()
    errs = []

    def check(name, ok, detail=""):
        if not ok:
            errs.append(f"rank {rank}: {name} FAILED {detail}")

    outdir = os.path.join(tempfile.gettempdir(), f"nnpm_ngpus_{os.environ.get('MASTER_PORT', '0')}")
    N = 16
    c = dict(cf.DEFAULTS, TopgridDimensions=[N, N, N], ParticleDimensions=[N, N, N], ProblemDefinition="UniformParticles",
             Cosmology="cosmology(h=0.7,OmegaM=0.3,OmegaR=0.0)", InitialRedshift=50.0, FinalRedshift=20.0, InitialDt=1e-4,
             CourantFactor=0.5, MaximumDt=1.0, DtFractionOfTime=0.1, StopCycle=6, OutputEveryNCycle=3, OutputDirectory=outdir,
             NGPUs=P, PartCapacityFactor=4.0)
    gp, gd, p = {}, {}, {}
    PM.InitializeParticleSimulation(gp, gd, p, c)
    rng = np.random.default_rng(7)                      # perturb the lattice identically on every rank
    p["x"] += 0.3 / N * rng.standard_normal(p["x"].shape)
    p["v"] += 1e-3 * rng.standard_normal(p["v"].shape)
    x0, v0 = p["x"].copy(), p["v"].copy()

    PM.DevicePM = make_adapter()
    log = []
    PM.evolvePMCosmology(gp, gd, p, c, log=log)

    oc = O.Cosmo(h=0.7, OmegaM=0.3, OmegaR=0.0)
    rc = dict(StartTime=c["StartTime"], StopTime=c["StopTime"], StartCycle=0, StopCycle=6, InitialDt=1e-4, CourantFactor=0.5,
              MaximumDt=1.0, DtFractionOfTime=0.1, InitialRedshift=50.0)
    xr, vr = x0.copy(), v0.copy()
    logr = []
    _, _, f = O.evolvePMCosmology(xr, vr, p["m"], (N, N, N), rc, oc, log=logr)
    check("steps", len(log) == len(logr) == 6)
    check("dt sequence", all(abs(a[1] - b[1]) <= 1e-12 * b[1] for a, b in zip(log, logr)), str([(a[1], b[1]) for a, b in zip(log, logr)]))
    check("x (reference order, every rank)", np.max(np.abs(p["x"] - xr)) < 1e-11, str(np.max(np.abs(p["x"] - xr))))
    check("v", np.max(np.abs(p["v"] - vr)) < 1e-11 * max(1.0, np.max(np.abs(vr))))
    check("phi gathered", np.max(np.abs(gd[1].d["Φ"] - f.phi)) <= 1e-11 * np.max(np.abs(f.phi)))
    check("rho gathered (density of the last step's deposit)", abs(gd[1].d["ρD"].sum() - p["m"] * N ** 3) < 1e-9 * p["m"] * N ** 3
          and np.max(np.abs(gd[1].d["ρD"] - f.rho)) <= 1e-11 * np.max(np.abs(f.rho)))
    dist.barrier()
    files = sorted(os.listdir(outdir)) if os.path.isdir(outdir) else []
    if rank == 0:
        check("rank 0 wrote the outputs", len(files) >= 2, str(files))
        import shutil
        shutil.rmtree(outdir, ignore_errors=True)
    check("output counter identical on every rank", c["CurrentOutputNumber"] >= 2)

    all_errs = [None] * P
    dist.all_gather_object(all_errs, errs)
    flat = [e for l in all_errs for e in l]
    if rank == 0:
        print("HOST_NGPUS_OK" if not flat else "HOST_NGPUS_FAIL\n" + "\n".join(flat))
    dist.destroy_process_group()
    return 1 if flat else 0


if __name__ == "__main__":
    sys.exit(main())
