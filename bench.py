#!/usr/bin/env python
"""bench.py -- particle-updates/s of one comoving PM step (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--n 1024] [--mesh 1024] [--ic grf|modes|clustered]

N = 1: the configuration the metric is quoted on, 1024^3 particles on a 1024^3 mesh (configs[3] of
BASELINE.json fits one 180 GB B200), seeded Gaussian-random-field Zel'dovich ICs generated on the device
(PowerSpectrumParticles + scaleFreePS(-1, .), RMS displacement 0.3 cell).
N > 1 (under torchrun): the same global problem slab-decomposed over N ranks ("strong" scaling --
the metric names a fixed 1024^3/1024^3 problem at 1/2/4/8 GPUs).

One JSON line on stdout (rank 0).
  value      device-resident steps (state in HBM before the timed region), CUDA events on the launching stream,
             max over ranks.
  e2e        the same step through the C ABI with HOST buffers -- upload x,v (pinned) -> step -> download x,v + stats
             inside the timed region (PCIe-bound by construction: the product copies only at outputs).
  roofline   the dominant kernel (fused gather + push): algorithmic bytes / CUDA-event time; `per_kernel` gives the
             same for every phase of the step (deposit, x / y / z transform passes, push), `step_*` the whole step
             against T = (Np B_p + Nc B_c) / (P BW_hbm) + NVLink bytes / 900 GB/s.
  parity     checked in this very run, outside the timed region: (a) one comoving step of a 64^3 problem on the SAME
             rank layout against the CPU oracle (phi / x / v), (b) at full size after the timed steps: mass
             conservation of the deposit, P(k) from nnpm_power_spectrum against the value stored by a 1-GPU run of
             the same seed and step count (profiles/pk_ref.json), halo violations.
  cpu_baseline / --impl reference   the C restatement of the reference's loops (oracle/pm_ref.c + pocketfft; the Julia
             reference cannot run in this image) on the host cores, on a bounded sample of the workload whose size the
             line states.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "particle-updates/sec per PM step (1024³ part, 1024³ mesh) at 1/2/4/8 B200; % HBM roofline"
UNIT = "particle-updates/s"
B_P_COMOVING = 144.0   # algorithmic bytes / particle / step
B_C = 96.0             # algorithmic bytes / mesh cell / step
# the same bytes split by phase: (per particle, per mesh cell)
PHASE_BYTES = {"deposit": (48.0, 8.0), "fft_x": (0.0, 32.0), "fft_y": (0.0, 32.0), "fft_z": (0.0, 16.0), "push": (96.0, 8.0)}
NVLINK_NOMINAL = 900e9   # per direction per GPU
NVLINK_MEASURED = 770e9  # peer copy measured on this pool (B200_PROFILING.md)
# hubble_time_gyr: the workload's time unit is pinned to the value the stored full-size P(k) references
# (profiles/pk_ref.json, recorded on 1 GPU in rounds 1-2) were produced with; the package default is now 9.77813952
# (noname_b200/cosmology.py).  A property of the synthetic workload, irrelevant to its cost.
COSMO = dict(h=0.7, OmegaM=0.3, OmegaR=0.0, hubble_time_gyr=9.77793)
ZINIT = 50.0
PK_REF = os.path.join(ROOT, "profiles", "pk_ref.json")


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def bind_to_gpu_numa(index):
    """N > 1: keep this rank's host thread -- and with it the pages of the pinned e2e buffers it allocates -- on
    the CPUs local to its GPU (sysfs local_cpulist of the GPU's PCI function).  Any failure is a no-op."""
    try:
        import torch
        p = torch.cuda.get_device_properties(index)
        bdf = f"{p.pci_domain_id:04x}:{p.pci_bus_id:02x}:{p.pci_device_id:02x}.0"
        with open(f"/sys/bus/pci/devices/{bdf}/local_cpulist") as f:
            txt = f.read().strip()
        cpus = set()
        for part in txt.split(","):
            if part:
                a, _, b = part.partition("-")
                cpus.update(range(int(a), int(b or a) + 1))
        allowed = os.sched_getaffinity(0)
        cpus &= allowed
        if cpus and len(cpus) < len(allowed):
            os.sched_setaffinity(0, cpus)
            return f"{bdf}: {len(cpus)} of {len(allowed)} CPUs"
        return f"{bdf}: no narrower CPU set ({txt or 'empty'})"
    except Exception as ex:  # noqa: BLE001 -- placement is an optimisation, never a failure
        return f"not bound ({ex!r})"


class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def stop(self, t0=None, t1=None):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, smax, reasons = [], None, set()
        for ts, r in self.rows:
            if t0 is not None and not (t0 - 0.05 <= ts <= t1 + 0.15):
                continue
            f = [q.strip() for q in r.split(",")]
            try:
                sm.append(float(f[0]))
                smax = float(f[1])
                for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                    if val.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": smax, "reasons": sorted(reasons),
                "samples": len(sm)}


def scalars(t, dt):
    """a(t+dt/4), a(t+dt/2), adot(t+dt/2), a(t+3dt/4) -- host cosmology, as evolvePMCosmology does."""
    from noname_b200 import cosmology as cos
    c = scalars.cosmo
    a1 = cos.a_from_t_internal(c, t + dt / 2 / 2, ZINIT, zwherea1=ZINIT)
    ak = cos.a_from_t_internal(c, t + dt / 2, ZINIT, zwherea1=ZINIT)
    adk = cos.dadt(c, ak, zwherea1=ZINIT)
    a2 = cos.a_from_t_internal(c, t + 3.0 * dt / 2 / 2, ZINIT, zwherea1=ZINIT)
    return a1, ak, adk, a2


def run_cpu(n, mesh, steps, warmup, threads):
    """The reference's loops (C restatement + pocketfft) on the host: comoving steps on an n^3 / mesh^3
    sample of the same synthetic workload.  `threads` is a count, or a list of candidate counts: each candidate then runs
    two probe steps on the evolving state and the fastest one does the timed steps.  Returns (pu/s, seconds per step,
    threads used, {candidate: probe seconds})."""
    from oracle import pm_ref as O
    from oracle.pm_ref_c import PMRefC
    from noname_b200 import cosmology as cos
    scalars.cosmo = cos.Cosmology(**COSMO)
    oc = O.Cosmo(**COSMO)
    ts, _ = O.start_stop_times(oc, ZINIT, 0.0)
    a0 = O.a_from_t_internal(oc, ts, ZINIT, zwherea1=ZINIT)
    ad0 = O.dadt_from_t_internal(oc, ts, ZINIT)
    x, v = O.synthetic_grf((n, n, n), 3, 12345, -1.0, 0.3 / mesh, ad0 / a0)
    m = (mesh ** 3) / float(n ** 3) * COSMO["OmegaM"]
    f = O.Fields((mesh, mesh, mesh), n ** 3, 3)
    state = {"t": ts, "dt": 1e-4}

    def step(ref):
        sc = scalars(state["t"], state["dt"])
        t0 = time.perf_counter()
        mabs, _ = ref.step_comoving(f, x, v, m, state["dt"], *sc)
        el = time.perf_counter() - t0
        state["t"] += state["dt"]
        state["dt"] = min(0.5 * (1.0 / mesh) / mabs, 0.1 * state["t"], 1.0)
        return el
    probes = {}
    if isinstance(threads, (list, tuple)):
        cands = sorted(set(int(q) for q in threads))
        if len(cands) > 1:
            for th in cands:
                ref = PMRefC(threads=th)
                step(ref)
                probes[th] = step(ref)
            threads = min(probes, key=probes.get)
        else:
            threads = cands[0]
    ref = PMRefC(threads=threads)
    el = [step(ref) for _ in range(warmup + steps)]
    sec = sum(el[warmup:]) / steps
    return n ** 3 / sec, sec, threads, probes


def parity_small_step(rank, world, local_rank, dist, opts=()):
    """CHECKER (oracle/ used as the checker only, outside every timed region): one comoving step of a 64^3 / 64^3
    problem on the same rank layout as the bench -- upload each rank's share, redistribute, step -- against the CPU
    oracle.  Returns normalised L-inf errors of phi and v and the max |dx| over all ranks."""
    import numpy as np
    import torch
    from noname_b200 import DevicePM
    from oracle import pm_ref as O
    N = 64
    dims, npart = (N, N, N), N ** 3
    x, v = O.synthetic_zeldovich(dims, 3, 2024, 16, 4, 0.3 / N, 0.2)
    m = 0.3
    f = O.Fields(dims, npart, 3)
    xr, vr = x.copy(), v.copy()
    sc = (1.001, 1.002, 0.81, 1.003)
    O.step_comoving(f, xr, vr, m, 0.01, *sc)
    dev = DevicePM(dims, npart, rank=3, device=local_rank, nranks=world, rank_id=rank, sort_every=1, part_capacity=npart)
    for kv in opts:   # the same nnpm_set_option overrides as the timed run
        k, _, v_ = kv.partition("=")
        dev.set_option(k, float(v_))
    try:
        if world > 1:
            uid = [DevicePM.comm_unique_id() if rank == 0 else None]
            dist.broadcast_object_list(uid, src=0)
            dev.comm_init(uid[0])
            lo, hi = rank * npart // world, (rank + 1) * npart // world
            dev.upload(np.ascontiguousarray(x[lo:hi]), np.ascontiguousarray(v[lo:hi]), m, first_id=lo)
            dev.redistribute()
        else:
            dev.upload(x, v, m)
        st = dev.step_comoving(0.01, *sc)
        nzl = N // world
        phi = dev.get_field("phi")
        xs, vs, ids = dev.download(ids=True)
        ref_phi = f.phi[rank * nzl:(rank + 1) * nzl]
        e = [float(np.max(np.abs(phi - ref_phi)) / np.max(np.abs(f.phi))),
             float(np.max(np.abs(xs - xr[ids]))) if len(ids) else 0.0,
             float(np.max(np.abs(vs - vr[ids])) / np.max(np.abs(vr))) if len(ids) else 0.0,
             float(st.n_halo_violations)]
    finally:
        dev.close()
    if world > 1:
        t = torch.tensor(e, dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e = [float(q) for q in t.tolist()]
    tol = 1e-12
    return {"what": f"one comoving step, 64^3 particles / 64^3 mesh, {world} rank(s), tile-sorted, vs CPU oracle",
            "phi_nlinf": e[0], "x_maxabs": e[1], "v_nlinf": e[2], "halo_violations": int(e[3]), "tol": tol,
            "ok": bool(e[0] < tol and e[1] < tol and e[2] < tol and e[3] == 0)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=16)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--n", "--particles-per-dim", dest="n", type=int, default=1024,
                    help="particles per dimension (under torchrun use the long spelling: --n is an ambiguous prefix of its own options)")
    ap.add_argument("--mesh", type=int, default=1024, help="mesh cells per dimension")
    ap.add_argument("--e2e-steps", type=int, default=2)
    ap.add_argument("--cpu-sample", type=int, default=192, help="n of the n^3/n^3 CPU-baseline sample (0 = skip)")
    ap.add_argument("--cpu-threads", type=int, default=0,
                    help="--impl reference: host threads of the C restatement (capped by the cores available); 0 = try 16, 32 and all "
                         "cores with two probe steps each and time the fastest (round 1: 32 threads were slower than 16; the "
                         "deposit and the reductions have since been made to scale, oracle/pm_ref_c.py)")
    ap.add_argument("--sort-every", type=int, default=16,
                    help="counting-sort cadence in steps; the default K = 16 timed steps contain exactly one sort")
    ap.add_argument("--deposit-mode", default="atomic")
    ap.add_argument("--opt", action="append", default=[], help="name=value passed to nnpm_set_option")
    ap.add_argument("--ic", default="grf", choices=["grf", "modes", "clustered"],
                    help="grf = seeded Gaussian random field (scaleFreePS n = -1), modes = 32 plane waves (round-1 workload), "
                         "clustered = BASELINE config 4 stand-in: half the particles in 10^5 Gaussian blobs (sigma 1.5 cells)")
    ap.add_argument("--no-parity", action="store_true", help="skip the in-run parity block")
    ap.add_argument("--pk-ref-write", action="store_true", help="N = 1: store this run's P(k) as the reference for its (n, mesh, ic, steps)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    W = max(args.warmup, 3)
    K = args.steps
    n, mesh = args.n, args.mesh
    ic_text = {"grf": "seeded Gaussian-random-field Zel'dovich ICs (scaleFreePS n = -1, RMS displacement 0.3 cell)",
               "modes": "synthetic Zel'dovich ICs (32 seeded plane waves, RMS displacement 0.3 cell)",
               "clustered": "synthetic clustered late-time state (50% of the particles in 10^5 Gaussian blobs, sigma = 1.5 cells, b^-1/2 mass function)"}[args.ic]
    workload = f"{n}^3 particles / {mesh}^3 mesh comoving PM step (LCDM Om=0.3 h=0.7), {ic_text}"
    config = {"workload": workload, "particles": n ** 3, "mesh_cells": mesh ** 3, "interpolation": "cic/cic", "ic": args.ic,
              "l2": "inputs (>= 48 B/particle resident state) exceed the 126 MB L2 for n >= 256; no flush needed",
              "sort_every": args.sort_every, "deposit_mode": args.deposit_mode,
              "parallelism": f"slab{world}" if world > 1 else "single",
              "cosmology": dict(COSMO, zinit=ZINIT)}

    if args.impl == "reference":
        if rank != 0:
            return 0
        ncores = len(os.sched_getaffinity(0))
        cand = [max(1, min(args.cpu_threads, ncores))] if args.cpu_threads > 0 else sorted({min(16, ncores), min(32, ncores), ncores})
        cs = args.cpu_sample or 192
        val, sec, threads, probes = run_cpu(cs, cs, K, W, cand)
        sample = (f"{cs}^3 particles / {cs}^3 mesh comoving steps of the same seeded random-field workload; C restatement of "
                  f"ParticleMesh.jl loops + pocketfft on {threads} of {ncores} host threads"
                  + (f" (fastest of the probed counts, s/step: {({k: round(q, 4) for k, q in probes.items()})})" if probes else "")
                  + " (the Julia reference, itself single-threaded, cannot run here)")
        # the config of this line describes what was RUN: the bounded sample, not the 1024^3 problem it stands in for
        rconfig = dict(config, workload=f"{cs}^3 particles / {cs}^3 mesh comoving PM step (LCDM Om=0.3 h=0.7), {ic_text} "
                                        f"-- bounded CPU sample of the {n}^3 / {mesh}^3 workload",
                       particles=cs ** 3, mesh_cells=cs ** 3, sample_of={"particles": n ** 3, "mesh_cells": mesh ** 3},
                       parallelism=f"{threads} host threads")
        line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": K,
                "warmup": W, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "strong",
                "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": rconfig,
                "cpu_baseline": {"value": val, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
                "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return 0

    import numpy as np
    import torch
    import torch.distributed as dist
    from noname_b200 import DevicePM, cosmology as cos, pinned_empty, pinned_free
    scalars.cosmo = cos.Cosmology(**COSMO)
    torch.cuda.set_device(local_rank)
    if world > 1:
        numa = bind_to_gpu_numa(local_rank)
        if os.environ.get("NNPM_VERBOSE"):
            print(f"[bench] rank {rank}: host placement {numa}", file=sys.stderr)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    stream = torch.cuda.current_stream()
    peak, peak_src = peaks()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t)
        return float(t.item())

    parity = None
    if not args.no_parity:
        try:
            parity = {"small_step": parity_small_step(rank, world, local_rank, dist, args.opt)}
        except Exception as ex:  # report, do not hide
            parity = {"small_step": {"ok": False, "error": repr(ex)}}

    dev = DevicePM((mesh, mesh, mesh), n ** 3, rank=3, device=local_rank, nranks=world, rank_id=rank, timing=True,
                   sort_every=args.sort_every, deposit_mode=args.deposit_mode, stream=stream.cuda_stream)
    for kv in args.opt:
        k, _, v = kv.partition("=")
        dev.set_option(k, float(v))
    if world > 1:
        uid = [DevicePM.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        dev.comm_init(uid[0])
    ts = scalars.cosmo.age_gyr(ZINIT) * (cos.gyr_in_s / cos.get_units(scalars.cosmo, ZINIT, zinit=ZINIT).T)
    a0 = cos.a_from_t_internal(scalars.cosmo, ts, ZINIT, zwherea1=ZINIT)
    ad0 = cos.dadt(scalars.cosmo, a0, zwherea1=ZINIT)
    m = (mesh ** 3) / float(n ** 3) * COSMO["OmegaM"]
    if args.ic == "clustered":
        dev.init_synthetic((n, n, n), kind="clustered", seed=777, nmodes=100, kmax=1000, disp_rms=1.5 / mesh, vfac=2e-5, m=m)
    elif args.ic == "modes":
        dev.init_synthetic((n, n, n), seed=12345, nmodes=32, kmax=8, disp_rms=0.3 / mesh, vfac=ad0 / a0, m=m)
    else:
        dev.init_grf((n, n, n), seed=12345, ps_index=-1.0, disp_rms=0.3 / mesh, vfac=ad0 / a0, m=m)

    state = {"t": ts, "dt": 1e-4}

    def one_step():
        sc = scalars(state["t"], state["dt"])
        st = dev.step_comoving(state["dt"], *sc)
        if world > 1:
            dev.reduce_stats(st)
        state["t"] += state["dt"]
        state["dt"] = min(0.5 * (1.0 / mesh) / (st.max_abs_v + 1e-30), 0.1 * state["t"], 1.0)
        return st

    halo = 0
    for _ in range(W):
        halo += one_step().n_halo_violations
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    wall0 = time.time()
    e0.record(stream)
    launches = 0
    ms_phase = {}
    untiled = 0
    for _ in range(K):
        st = one_step()
        launches += st.kernel_launches
        halo += st.n_halo_violations
        untiled = max(untiled, st.n_untiled)
        for k, v in st.as_dict()["ms"].items():
            ms_phase[k] = ms_phase.get(k, 0.0) + v
    e1.record(stream)
    barrier()
    wall1 = time.time()
    ms_total = max_over_ranks(e0.elapsed_time(e1))
    clocks = sampler.stop(wall0, wall1) if rank == 0 else None
    ms_step = ms_total / K
    value = n ** 3 / (ms_step * 1e-3)
    ms_phase = {k: max_over_ranks(v / K) for k, v in ms_phase.items()}   # reduce_stats already made them global maxima
    np_local = dev.local_count
    np_max = max_over_ranks(float(np_local))

    # roofline of the dominant kernel (fused gather + kick + drift push): algorithmic bytes / event time
    push_ms = ms_phase["push"]
    push_bytes = PHASE_BYTES["push"][0] * np_max
    ach = push_bytes / (push_ms * 1e-3) / 1e9
    traffic = None
    tp = os.path.join(ROOT, "profiles", "push_traffic.json")
    if os.path.exists(tp):
        try:
            with open(tp) as f:
                tj = json.load(f)
            if tj.get("particles") == np_local:
                traffic = tj.get("dram_bytes_per_launch")
        except Exception:
            pass
    per_kernel = {}
    for ph, (bp, bc) in PHASE_BYTES.items():
        t_ms = ms_phase.get(ph, 0.0)
        if t_ms > 0:
            b = (bp * n ** 3 + bc * mesh ** 3) / world
            per_kernel[ph] = {"ms": t_ms, "algorithmic_bytes": b, "achieved_gbs": b / (t_ms * 1e-3) / 1e9,
                              "frac": b / (t_ms * 1e-3) / 1e9 / peak}
    roofline = {"bound": "hbm", "kernel": "k_push_tile (TMA-staged phi box; gradient + CIC gather + comoving kick + drift + wrap + max reductions)",
                "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "traffic": traffic,
                "peak_source": peak_src, "algorithmic_bytes_per_launch": push_bytes,
                "kernel_ms": push_ms, "phase_ms": ms_phase, "per_kernel": per_kernel,
                "per_kernel_note": "phase times are CUDA events inside the step (max over ranks); deposit includes the mesh zero-fill and, "
                                   "with N > 1, the ghost-plane fold; fft_y (N > 1: forward pass) and the transposes are listed under comm"}
    t_hbm = ((n ** 3) * B_P_COMOVING + (mesh ** 3) * B_C) / world / (peak * 1e9)
    nvl = 2 * 16.0 * (mesh // 2 + 1) * mesh * mesh * (world - 1) / world ** 2 if world > 1 else 0.0
    t_roof = t_hbm + nvl / NVLINK_NOMINAL
    roofline["step_roofline_ms"] = t_roof * 1e3
    roofline["step_frac"] = t_roof * 1e3 / ms_step
    roofline["nvlink_gbs_assumed"] = NVLINK_NOMINAL / 1e9
    if world > 1:
        roofline["step_frac_at_measured_nvlink_770"] = (t_hbm + nvl / NVLINK_MEASURED) * 1e3 / ms_step
        roofline["nvlink_bytes_per_gpu_per_step"] = nvl
        roofline["comm_frac_of_step"] = ms_phase.get("comm", 0.0) / ms_step

    # parity at full size (outside the timed region): mass, P(k) against the stored 1-GPU value, halo violations
    if parity is not None:
        try:
            dev.deposit()
            rho = dev.get_field("rho")
            mass = sum_over_ranks(float(rho.sum(dtype=np.float64)))
            del rho
            want = m * float(n ** 3)
            parity["mass_rel"] = abs(mass - want) / want
            kk, pk = dev.power_spectrum()
            key = f"{n}-{mesh}-{args.ic}-{W + K}steps-{args.deposit_mode}"
            ref = {}
            if os.path.exists(PK_REF):
                with open(PK_REF) as f:
                    ref = json.load(f)
            if args.pk_ref_write and world == 1 and rank == 0:
                ref[key] = {"power": [float(q) for q in pk], "written_by": "bench.py --pk-ref-write, 1 GPU"}
                with open(PK_REF, "w") as f:
                    json.dump(ref, f)
            if key in ref:
                pr = np.asarray(ref[key]["power"])
                ok = np.isfinite(pr) & (pr != 0)
                parity["pk_vs_P1_maxrel"] = float(np.max(np.abs(pk[ok] / pr[ok] - 1.0)))
            else:
                parity["pk_vs_P1_maxrel"] = None
                parity["pk_note"] = f"no stored 1-GPU P(k) for {key} (profiles/pk_ref.json)"
            parity["n_halo_violations"] = int(sum_over_ranks(float(halo)))
            parity["max_untiled_per_step"] = int(max_over_ranks(float(untiled)))
            parity["ok"] = bool(parity["small_step"].get("ok") and parity["mass_rel"] < 1e-12 and parity["n_halo_violations"] == 0
                                and (parity["pk_vs_P1_maxrel"] is None or parity["pk_vs_P1_maxrel"] < 1e-6))
        except Exception as ex:  # report, do not hide
            parity["ok"] = False
            parity["error"] = repr(ex)

    # e2e: host buffers through the C ABI: upload -> step -> download, every step
    e2e = None
    if args.e2e_steps > 0:
        try:
            nl = dev.local_count
            # pinned host buffers sized once for the per-rank capacity: with N > 1 migration changes the
            # local count every step and page-locking tens of GB inside the timed region would dominate it
            cap = nl if world == 1 else int(nl * 1.05) + 65536
            bx, bv = pinned_empty((cap, 3)), pinned_empty((cap, 3))
            hx, hv = bx[:nl], bv[:nl]
            dev.download(hx, hv)
            barrier()
            tE, h2d, d2h = [], 0, 0
            for it in range(1 + args.e2e_steps):
                barrier()
                t0 = time.perf_counter()
                dev.upload(hx, hv, m)
                sc = scalars(state["t"], state["dt"])
                st = dev.step_comoving(state["dt"], *sc)
                if world > 1:
                    dev.reduce_stats(st)
                n2 = dev.local_count
                if n2 > cap:
                    raise RuntimeError(f"rank {rank}: {n2} particles exceed the host buffer capacity {cap}")
                h2d, d2h = 2 * nl * 24, 2 * n2 * 24 + 32
                nl = n2
                hx, hv = bx[:nl], bv[:nl]
                dev.download(hx, hv)
                torch.cuda.synchronize()
                tE.append(max_over_ranks(time.perf_counter() - t0))
            sec = sum(tE[1:]) / len(tE[1:])
            tot = torch.tensor([float(h2d), float(d2h)], dtype=torch.float64, device="cuda")
            if world > 1:
                dist.all_reduce(tot)
            e2e = {"value": n ** 3 / sec, "unit": UNIT, "h2d_bytes_per_step": int(tot[0].item()),
                   "d2h_bytes_per_step": int(tot[1].item()), "ms_per_step": sec * 1e3,
                   "steps": len(tE) - 1, "api": "nnpm_upload_particles -> nnpm_step_comoving -> nnpm_download_particles (pinned host buffers)",
                   "note": "every step re-uploads and re-downloads the whole particle state (the reference keeps it in host arrays); "
                           "the product's own loop copies only at outputs"}
            pinned_free(bx), pinned_free(bv)
        except Exception as ex:  # report, do not hide
            e2e = {"value": None, "unit": UNIT, "error": repr(ex)}
    dev.close()

    cpu = None
    if rank == 0 and args.cpu_sample:
        cs = args.cpu_sample
        val, sec, _, _ = run_cpu(cs, cs, 2, 1, 1)
        cpu = {"value": val, "unit": UNIT, "cores": 1, "kind": "port",
               "sample": f"2 comoving steps of a {cs}^3 / {cs}^3 sample of the same synthetic workload ({sec:.2f} s/step); "
                         "C restatement of ParticleMesh.jl + pocketfft, 1 thread like the reference (Readme.md:53); "
                         "the Julia reference itself cannot run in this image"}
    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
                "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
                "dtype": "f64", "data": "synthetic", "config": config, "clocks": clocks, "e2e": e2e,
                "gpu_launches": int(launches), "roofline": roofline, "parity": parity, "cpu_baseline": cpu}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
