"""world_size-2 gloo test (CPU) of the N>1 host logic: shard ranges, scatter of QP records from rank 0, per-rank
solve (with the CPU oracle standing in for the device, this is a test of the plumbing), gather in batch order."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from acados_b200.sharding import shard_range


def test_shard_range_partitions():
    for nb in (0, 1, 7, 4096, 65537):
        for w in (1, 2, 3, 8):
            cuts = [shard_range(nb, r, w) for r in range(w)]
            assert cuts[0][0] == 0 and cuts[-1][1] == nb
            assert all(cuts[i][1] == cuts[i + 1][0] for i in range(w - 1))
            sizes = [b - a for a, b in cuts]
            assert max(sizes) - min(sizes) <= 1


def _worker(rank, world, port, tmp):
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, root)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from acados_b200 import problems as P
    from acados_b200.binding import default_opts
    from acados_b200.sharding import gather_records, scatter_records
    from oracle import oracle_binding as ob
    nb = 7
    b = P.mass_spring(nb, seed=5, x0_scale=0.5)
    mine = scatter_records(torch.from_numpy(b.qp) if rank == 0 else None)
    lo, hi = shard_range(nb, rank, world)
    assert mine.shape == (hi - lo, b.layout.qp_stride)
    assert np.array_equal(mine.numpy(), b.qp[lo:hi])
    sub = P.Batch(b.shape, b.layout, np.ascontiguousarray(mine.numpy()))
    sol, info = ob.oracle_solve(sub, default_opts(), nthreads=1)
    t = torch.tensor([float(rank + 1)])
    dist.all_reduce(t, op=dist.ReduceOp.MAX)          # the bench's max-over-ranks timing reduction
    assert t.item() == world
    allsol = gather_records(torch.from_numpy(sol), nb)
    if rank == 0:
        full, _ = ob.oracle_solve(b, default_opts(), nthreads=1)
        assert np.array_equal(allsol.numpy(), full)
        open(os.path.join(tmp, "ok"), "w").write("ok")
    dist.destroy_process_group()


def test_scatter_solve_gather_world2(built, tmp_path):
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert (tmp_path / "ok").exists()
