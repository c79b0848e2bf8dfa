"""Host-side plumbing of the slab-decomposed run (conf key NGPUs > 1): one process per GPU, launched with
torchrun / mpirun-style environment variables (RANK, WORLD_SIZE, LOCAL_RANK, MASTER_ADDR, MASTER_PORT).

The reference has no parallelism (Readme.md:53); its host state -- gp, gd, p -- stays what it was: every
rank holds the full dictionaries, the device contexts hold one slab each, and at outputs the slabs are
gathered back so that gd / p are complete and identical on every rank (rank 0 writes the files).
torch.distributed is used for the rendezvous and the host-side gathers only (gloo on CPU tensors); the
data path between GPUs lives in libnnpm (NCCL + NVLink peer memory).
"""
from __future__ import annotations

import os

import numpy as np


def world():
    """(rank, world_size, local_rank) from the launcher's environment; (0, 1, 0) when not launched in parallel."""
    return (int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")),
            int(os.environ.get("LOCAL_RANK", os.environ.get("RANK", "0"))))


def _dist():
    import torch.distributed as dist
    if not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29400")
        dist.init_process_group("gloo")
    return dist


def require_world(ngpus):
    rank, size, local = world()
    if size != ngpus:
        raise RuntimeError(f"NGPUs = {ngpus} needs {ngpus} processes (one per GPU, e.g. `torchrun --nproc-per-node {ngpus}`); "
                           f"this process sees WORLD_SIZE = {size}")
    _dist()
    return rank, size, local


def broadcast_bytes(b, src=0):
    box = [b]
    _dist().broadcast_object_list(box, src=src)
    return box[0]


def allgather_rows(a):
    """concatenate per-rank arrays of different lengths along axis 0 (host)."""
    import torch
    dist = _dist()
    n = torch.tensor([a.shape[0]], dtype=torch.int64)
    counts = [torch.zeros(1, dtype=torch.int64) for _ in range(dist.get_world_size())]
    dist.all_gather(counts, n)
    counts = [int(c.item()) for c in counts]
    mx = max(counts)
    pad = np.zeros((mx,) + a.shape[1:], dtype=a.dtype)
    pad[:a.shape[0]] = a
    outs = [torch.zeros(pad.shape, dtype=torch.from_numpy(pad).dtype) for _ in counts]
    dist.all_gather(outs, torch.from_numpy(pad))
    return np.concatenate([o.numpy()[:c] for o, c in zip(outs, counts)], axis=0)


def allgather_slabs(a):
    """equal-size slabs (N3/P, N2, N1) -> the full (N3, N2, N1) array on every rank."""
    import torch
    dist = _dist()
    outs = [torch.zeros(a.shape, dtype=torch.float64) for _ in range(dist.get_world_size())]
    dist.all_gather(outs, torch.from_numpy(np.ascontiguousarray(a)))
    return np.concatenate([o.numpy() for o in outs], axis=0)
