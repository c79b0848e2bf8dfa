"""CPU ORACLE (test infrastructure, not product code) -- the slab-decomposed form of the PM step,
in numpy over torch.distributed (gloo on CPU), one process per slab.

The reference has no parallelism (Readme.md:53); the decomposition is this repo's design
(DESIGN.md "Multi-GPU"), and this file restates the *protocol* the CUDA library follows in
noname_b200/csrc/nnpm_comm.cuh so that it can be exercised without GPUs:

    mesh      rank r owns planes k in [r*N3/P, (r+1)*N3/P); local array of nzl+5 planes
              (2 ghosts below, 3 above), local plane p <-> global plane kz0 + p - 2
    deposit   local scatter (particles may sit one plane outside the slab after the half-drift),
              then fold: local plane 1 -> lower neighbour's top owned plane, local planes
              nzl+2, nzl+3 -> upper neighbour's first two owned planes           (p2p)
    solve     2-D rfft per owned plane -> all-to-all slab transpose -> z FFT, Green's function
              (ParticleMesh.jl:482-502), z inverse -> all-to-all back -> 2-D irfft
    phi halo  2 planes from below, 3 from above                                  (p2p)
    push      gradient + CIC gather + kick + drift from the local planes (ParticleMesh.jl:631-728,
              800-807), then particles whose plane left the slab migrate          (all-to-all)

Per-particle arithmetic is oracle/pm_ref.py's; results must equal the single-process oracle up
to summation order (rho) and FFT rounding (phi).  PARITY UNPINNED (see oracle/pm_ref.py).
Only tests/ import this module.
"""
from __future__ import annotations

import math

import numpy as np
import scipy.fft as sfft
import torch
import torch.distributed as dist

from . import pm_ref as O

GLO, GHI = 2, 3


def _t(a):
    return torch.from_numpy(np.ascontiguousarray(a))


def sendrecv(send_to, send_arr, recv_from, recv_shape, dtype=np.float64):
    """one paired exchange (isend/irecv, deadlock-free for rings of any size, incl. P == 2)"""
    out = torch.empty(recv_shape, dtype=torch.from_numpy(np.empty(0, dtype=dtype)).dtype)
    reqs = [dist.isend(_t(send_arr), send_to), dist.irecv(out, recv_from)]
    for r in reqs:
        r.wait()
    return out.numpy()


def alltoall(send, shapes):
    """grouped isend/irecv to every peer (gloo has no all_to_all; this is also how the CUDA library
    issues its transposes: ncclGroupStart + P-1 ncclSend/ncclRecv pairs + a local copy).
    float64 arrays; shapes[s] = shape of what rank s sends to this rank."""
    P, r = dist.get_world_size(), dist.get_rank()
    outs = [torch.empty(tuple(shapes[s]), dtype=torch.float64) for s in range(P)]
    reqs = []
    for o in range(1, P):
        to, frm = (r + o) % P, (r - o) % P
        reqs.append(dist.isend(_t(send[to]), to))
        reqs.append(dist.irecv(outs[frm], frm))
    outs[r].copy_(_t(send[r]))
    for q in reqs:
        q.wait()
    return [o.numpy() for o in outs]


class SlabPM:
    def __init__(self, dims):
        self.P, self.r = dist.get_world_size(), dist.get_rank()
        self.n1, self.n2, self.n3 = dims
        assert self.n3 % self.P == 0 and self.n2 % self.P == 0 and self.n3 // self.P >= 4
        self.nzl = self.n3 // self.P
        self.njl = self.n2 // self.P
        self.kz0 = self.nzl * self.r
        self.up, self.dn = (self.r + 1) % self.P, (self.r - 1) % self.P
        self.mesh = np.zeros((self.nzl + GLO + GHI, self.n2, self.n1))

    # -- ownership ------------------------------------------------------------------------
    def owner(self, x):
        z = x[:, 2].copy()
        O.make_periodic(z)
        k = np.mod(np.floor(z * float(self.n3)).astype(np.int64), self.n3)
        return k // self.nzl

    def local_plane(self, k):
        d = k - self.kz0
        d = np.where(d >= self.nzl + GHI, d - self.n3, d)
        d = np.where(d < -GLO, d + self.n3, d)
        assert np.all((d >= -GLO) & (d < self.nzl + GHI)), "particle left the slab's ghost region"
        return d + GLO

    def migrate(self, x, v, ids):
        """route every particle to the rank owning floor(x3*N3): counts first, then payload"""
        own = self.owner(x)
        rec = np.concatenate([x, v, ids[:, None].astype(np.float64)], axis=1)
        send = [np.ascontiguousarray(rec[own == d]) for d in range(self.P)]
        cnt = torch.tensor([len(s) for s in send], dtype=torch.int64)
        allc = [torch.empty(self.P, dtype=torch.int64) for _ in range(self.P)]
        dist.all_gather(allc, cnt)                      # row s = counts sent by rank s (as nnpm_comm.cuh does)
        outs = alltoall(send, [(int(allc[s][self.r]), 7) for s in range(self.P)])
        got = np.concatenate(outs, axis=0)
        return got[:, :3].copy(), got[:, 3:6].copy(), got[:, 6].astype(np.int64)

    # -- deposit --------------------------------------------------------------------------
    def deposit(self, x, m):
        """x already wrapped into [0,1]; CIC (ParticleMesh.jl:383-405) into the local planes + fold"""
        self.mesh[...] = 0.0
        i, ip, dx = O._cell(x[:, 0], self.n1)
        j, jp, dy = O._cell(x[:, 1], self.n2)
        k, kp, dz = O._cell(x[:, 2], self.n3)
        p0, p1 = self.local_plane(k), self.local_plane(kp)
        n1, n2 = self.n1, self.n2

        def lin(a, b, c):
            return a + n1 * (b + n2 * c)
        idx = np.stack([lin(i, j, p0), lin(i, jp, p0), lin(ip, j, p0), lin(ip, jp, p0),
                        lin(i, j, p1), lin(i, jp, p1), lin(ip, j, p1), lin(ip, jp, p1)], axis=1)
        w = np.stack([m * (1 - dx) * (1 - dy) * (1 - dz), m * (1 - dx) * dy * (1 - dz),
                      m * dx * (1 - dy) * (1 - dz), m * dx * dy * (1 - dz),
                      m * (1 - dx) * (1 - dy) * dz, m * (1 - dx) * dy * dz,
                      m * dx * (1 - dy) * dz, m * dx * dy * dz], axis=1)
        np.add.at(self.mesh.reshape(-1), idx.reshape(-1), w.reshape(-1))
        assert not self.mesh[0].any() and not self.mesh[self.nzl + GLO + 2].any()
        nzl = self.nzl
        from_up = sendrecv(self.dn, self.mesh[1], self.up, (n2, n1))                  # up's "below" plane
        from_dn = sendrecv(self.up, self.mesh[GLO + nzl:GLO + nzl + 2], self.dn, (2, n2, n1))
        self.mesh[GLO + nzl - 1] += from_up
        self.mesh[GLO:GLO + 2] += from_dn
        return self.mesh[GLO:GLO + nzl]

    # -- solve ----------------------------------------------------------------------------
    def _transpose(self, a, fwd):
        """fwd: A[kl][j][ic] (all j) -> T[k][jl][ic] (all k); bwd: the inverse"""
        P, nzl, njl, icn = self.P, self.nzl, self.njl, a.shape[2]
        if fwd:
            send = [np.ascontiguousarray(a[:, d * njl:(d + 1) * njl, :]).view(np.float64) for d in range(P)]
            outs = alltoall(send, [(nzl, njl, 2 * icn)] * P)
            return np.concatenate([o.view(np.complex128) for o in outs], axis=0)
        send = [np.ascontiguousarray(a[d * nzl:(d + 1) * nzl]).view(np.float64) for d in range(P)]
        outs = alltoall(send, [(nzl, njl, 2 * icn)] * P)
        return np.concatenate([o.view(np.complex128) for o in outs], axis=1)

    def solve(self):
        """compute_potential (ParticleMesh.jl:473-505) on the owned planes; fills the phi halo"""
        n1, n2, n3, nzl = self.n1, self.n2, self.n3, self.nzl
        own = self.mesh[GLO:GLO + nzl]
        A = sfft.rfft2(own, axes=(1, 2))
        T = self._transpose(A, True)
        T = sfft.fft(T, axis=0)
        j0 = self.njl * self.r
        ii, jj, kk = np.arange(n1 // 2 + 1), np.arange(j0, j0 + self.njl), np.arange(n3)
        s1 = np.sin(2.0 * math.pi * ii / n1 / 2) ** 2
        s2 = np.sin(2.0 * math.pi * jj / n2 / 2) ** 2
        s3 = np.sin(2.0 * math.pi * kk / n3 / 2) ** 2
        ginv = -4.0 * (s1[None, None, :] + s2[None, :, None] + s3[:, None, None])
        nz = ginv != 0.0
        T[nz] = T[nz] / ginv[nz]
        if self.r == 0:
            T[0, 0, 0] = 0.0
        T = sfft.ifft(T, axis=0)
        A = self._transpose(T, False)
        own[...] = sfft.irfft2(A, s=(n2, n1), axes=(1, 2)) * 1 / float(n1) ** 2
        lo = sendrecv(self.up, own[nzl - 2:], self.dn, (2, n2, n1))      # my top two -> up's low ghosts
        hi = sendrecv(self.dn, own[:3], self.up, (3, n2, n1))            # my first three -> dn's high ghosts
        self.mesh[:GLO] = lo
        self.mesh[GLO + nzl:] = hi
        return own

    # -- gather ---------------------------------------------------------------------------
    def gather(self, x):
        """compute_acceleration (:631-637) on local planes 1 .. nzl+3, then cic_interp (:694-721)"""
        ph = self.mesh
        acc = np.zeros(ph.shape + (3,))
        acc[..., 0] = (np.roll(ph, -1, axis=2) - np.roll(ph, 1, axis=2)) / 2
        acc[..., 1] = (np.roll(ph, -1, axis=1) - np.roll(ph, 1, axis=1)) / 2
        acc[1:-1, :, :, 2] = (ph[2:] - ph[:-2]) / 2
        acc *= float(self.n1)
        np.negative(acc, out=acc)
        i, ip, dx = O._cell(x[:, 0], self.n1)
        j, jp, dy = O._cell(x[:, 1], self.n2)
        k, kp, dz = O._cell(x[:, 2], self.n3)
        k, kp = self.local_plane(k), self.local_plane(kp)
        assert np.all(k >= 1) and np.all(kp <= self.nzl + GLO + 1)
        pa = np.empty_like(x)
        for mm in range(3):
            am = acc[..., mm]
            pa[:, mm] = (am[k, j, i] * (1.0 - dx) * (1.0 - dy) * (1.0 - dz)
                         + am[k, j, ip] * dx * (1.0 - dy) * (1.0 - dz)
                         + am[k, jp, i] * (1.0 - dx) * dy * (1.0 - dz)
                         + am[k, jp, ip] * dx * dy * (1.0 - dz)
                         + am[kp, j, i] * (1.0 - dx) * (1.0 - dy) * dz
                         + am[kp, j, ip] * dx * (1.0 - dy) * dz
                         + am[kp, jp, i] * (1.0 - dx) * dy * dz
                         + am[kp, jp, ip] * dx * dy * dz)
        return pa

    # -- steps ----------------------------------------------------------------------------
    def step_comoving(self, x, v, ids, m, dt, a_d1, a_k, adot_k, a_d2):
        """ParticleMesh.jl:853-868 on this rank's particles; returns the migrated (x, v, ids)"""
        O.drift(x, v, dt / 2 / a_d1)
        O.make_periodic(x)
        self.deposit(x, m)
        self.solve()
        pa = self.gather(x)
        O.kick(v, pa, dt, a_k, adot_k)
        O.drift(x, v, dt / 2 / a_d2)
        O.make_periodic(x)
        return self.migrate(x, v, ids)

    def step_static(self, x, v, ids, m, dt):
        """ParticleMesh.jl:766-778"""
        O.make_periodic(x)
        self.deposit(x, m)
        self.solve()
        pa = self.gather(x)
        O.drift(x, v, dt / 2)
        O.kick(v, pa, dt)
        O.drift(x, v, dt / 2)
        return self.migrate(x, v, ids)

    def global_max(self, val):
        t = torch.tensor([val], dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())
