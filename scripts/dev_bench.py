"""Quick throughput probe on a GPU box: python scripts/dev_bench.py <cfg> <nbatch> [warps...]"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from acados_b200 import problems as P
from acados_b200.binding import CuipmSolver, default_opts, INFO_DTYPE
name = sys.argv[1]; nb = int(sys.argv[2]); Ws = [int(w) for w in sys.argv[3:]] or [1, 2, 4]
t0 = time.time(); b = P.named_config(name, nb); print("gen", time.time() - t0, "s; qp MB", b.qp.nbytes / 1e6, flush=True)
o = default_opts()
s = CuipmSolver(b.shape, nb)
dq = torch.from_numpy(b.qp).cuda()
dsol = torch.zeros((nb, b.layout.sol_stride), dtype=torch.float64, device="cuda")
dinfo = torch.zeros(nb * INFO_DTYPE.itemsize, dtype=torch.uint8, device="cuda")
for W in Ws:
    s.set_tuning("warps", W)
    for rep in range(3):
        s.solve_device(nb, dq.data_ptr(), dsol.data_ptr(), dinfo.data_ptr(), o, sync=True)
        ms = s.last_kernel_ms
    info = np.frombuffer(dinfo.cpu().numpy().tobytes(), dtype=INFO_DTYPE)
    print(f"{name} nb={nb} W={W}: kernel {ms:.2f} ms -> {nb/ms*1e3:.0f} QP/s; iters mean {info['iter'].mean():.2f} max {info['iter'].max()} "
          f"status hist {np.bincount(info['status'], minlength=4)} lq {info['lq_count'].sum()}", flush=True)
