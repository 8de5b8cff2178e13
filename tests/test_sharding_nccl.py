"""world_size-2 NCCL test (two GPUs of one box, -m gpu): the records of a batch held by rank 0 on its device are scattered
over NVLink, every rank solves its shard with the CUDA path, the solutions are gathered on rank 0 in batch order and compared
with a single-GPU solve of the whole batch (bit-identical: a QP's result does not depend on where it is solved).
Skipped on boxes with one GPU (run with ``gpurun --gpus 2``)."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def _worker(rank, world, port, tmp):
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, root)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    from acados_b200 import problems as P
    from acados_b200.binding import INFO_DTYPE, CuipmSolver, default_opts
    from acados_b200.sharding import gather_records, scatter_records, shard_range
    nb = 257          # odd: the two shards differ in size
    b = P.chain_mass(nb, N=20, seed=7)
    o = default_opts()
    recs = torch.from_numpy(b.qp).cuda() if rank == 0 else None
    mine = scatter_records(recs)
    lo, hi = shard_range(nb, rank, world)
    assert mine.is_cuda and mine.shape == (hi - lo, b.layout.qp_stride)
    assert np.array_equal(mine.cpu().numpy(), b.qp[lo:hi])
    s = CuipmSolver(b.shape, hi - lo, device=rank)
    d_sol = torch.zeros((hi - lo, b.layout.sol_stride), dtype=torch.float64, device="cuda")
    d_info = torch.zeros((hi - lo) * INFO_DTYPE.itemsize, dtype=torch.uint8, device="cuda")
    s.solve_device(hi - lo, mine.data_ptr(), d_sol.data_ptr(), d_info.data_ptr(), o, sync=True)
    allsol = gather_records(d_sol, nb)
    s.close()
    if rank == 0:
        s1 = CuipmSolver(b.shape, nb, device=0)
        full, info = s1.solve(b.qp, o)
        s1.close()
        assert (info["status"] == 0).all()
        assert np.array_equal(allsol.cpu().numpy(), full)
        open(os.path.join(tmp, "ok"), "w").write("ok")
    dist.destroy_process_group()


def test_nccl_scatter_solve_gather_world2(built, tmp_path):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert (tmp_path / "ok").exists()
