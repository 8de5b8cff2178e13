"""Batch sharding across ranks (one process per GPU).  The batch of independent QPs is the only sharding axis
(SURVEY.md section 8(e)); there is no exchange step inside the solve, so the only collectives are the optional
scatter of packed QP records from a root rank and the gather of packed solutions / summaries back to it.
Reference analogue: ``#pragma omp parallel for`` over capsules in the generated batch solver
(interfaces/acados_template/acados_template/c_templates_tera/acados_solver.in.c:3223-3243)."""
from __future__ import annotations

from typing import Optional, Tuple

import numpy as np


def shard_range(nbatch: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous, balanced slice [lo, hi) of the batch owned by ``rank`` (first ``nbatch % world`` ranks get one more)."""
    base, rem = divmod(nbatch, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def _comm_device(device=None):
    """Device the collectives run on: the caller's choice, else the current CUDA device under nccl (which moves CUDA
    tensors only) and the CPU under gloo."""
    import torch
    import torch.distributed as dist
    if device is not None:
        return torch.device(device)
    if dist.get_backend() == "nccl":
        return torch.device("cuda", torch.cuda.current_device())
    return torch.device("cpu")


def scatter_records(records, root: int = 0, device=None):
    """Root holds (nbatch, stride) records; every rank receives its shard_range slice.  Works on CPU tensors with
    gloo and CUDA tensors with nccl (send/recv based, sizes known from nbatch).  ``device``: where the shards live
    (default: the current CUDA device under nccl, the CPU under gloo); the root's records must be there already."""
    import torch
    import torch.distributed as dist
    rank, world = dist.get_rank(), dist.get_world_size()
    dev = _comm_device(device)
    if rank == root and records.device != dev:
        raise ValueError(f"scatter_records: the root's records are on {records.device}, the collectives run on {dev}")
    meta = torch.zeros(2, dtype=torch.int64, device=dev)
    if rank == root:
        meta[0], meta[1] = records.shape[0], records.shape[1]
    dist.broadcast(meta, root)
    nbatch, stride = int(meta[0]), int(meta[1])
    lo, hi = shard_range(nbatch, rank, world)
    if rank == root:
        reqs = []
        for r in range(world):
            if r == root:
                continue
            a, b = shard_range(nbatch, r, world)
            reqs.append(dist.isend(records[a:b].contiguous(), r))
        mine = records[lo:hi].clone()
        for q in reqs:
            q.wait()
        return mine
    mine = torch.empty((hi - lo, stride), dtype=torch.float64, device=dev)
    dist.recv(mine, root)
    return mine


def gather_records(mine, nbatch: int, root: int = 0):
    """Inverse of scatter_records: root receives every rank's shard in batch order (others get None)."""
    import torch
    import torch.distributed as dist
    rank, world = dist.get_rank(), dist.get_world_size()
    if rank != root:
        dist.send(mine.contiguous(), root)
        return None
    out = torch.empty((nbatch, mine.shape[1]), dtype=mine.dtype, device=mine.device)
    for r in range(world):
        a, b = shard_range(nbatch, r, world)
        if r == root:
            out[a:b] = mine
        else:
            dist.recv(out[a:b], r)
    return out
