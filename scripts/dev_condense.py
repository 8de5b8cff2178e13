"""Timing probe of the device block condensing on the quadrotor-sized shape (BASELINE config 4: nx=12 nu=4 N=50 -> N2=10):
python scripts/dev_condense.py [nbatch]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from acados_b200 import problems as P
from acados_b200.binding import CuipmCondenser, CuipmSolver, default_opts, INFO_DTYPE
nb = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
b = P.named_config("c4", nb)
o = default_opts()
dc = CuipmCondenser(b.shape, 10)
cl = dc.condensed_layout
print("condensed stage dims nu", dc.condensed_shape.nu[:2], "nx", dc.condensed_shape.nx[:2], "ng", dc.condensed_shape.ng[:2], "ns", dc.condensed_shape.ns[:2],
      "| record KB", b.layout.qp_stride * 8 // 1024, "->", cl.qp_stride * 8 // 1024, flush=True)
d_qp = torch.from_numpy(b.qp).cuda()
d_c = torch.empty((nb, cl.qp_stride), dtype=torch.float64, device="cuda")
d_sc = torch.zeros((nb, cl.sol_stride), dtype=torch.float64, device="cuda")
d_s = torch.zeros((nb, b.layout.sol_stride), dtype=torch.float64, device="cuda")
d_s_un = torch.zeros((nb, b.layout.sol_stride), dtype=torch.float64, device="cuda")
d_info = torch.zeros(nb * INFO_DTYPE.itemsize, dtype=torch.uint8, device="cuda")
su, sc = CuipmSolver(b.shape, nb), CuipmSolver(dc.condensed_shape, nb)
ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
for rep in range(3):
    torch.cuda.synchronize()
    ev[0].record(); dc.condense(nb, d_qp.data_ptr(), d_c.data_ptr()); ev[1].record()
    torch.cuda.synchronize()
    sc.solve_device(nb, d_c.data_ptr(), d_sc.data_ptr(), d_info.data_ptr(), o, sync=True); t_sc = sc.last_kernel_ms
    ic = np.frombuffer(d_info.cpu().numpy().tobytes(), dtype=INFO_DTYPE).copy()
    ev[2].record(); dc.expand(nb, d_qp.data_ptr(), d_sc.data_ptr(), d_s.data_ptr()); ev[3].record()
    torch.cuda.synchronize()
    su.solve_device(nb, d_qp.data_ptr(), d_s_un.data_ptr(), d_info.data_ptr(), o, sync=True); t_su = su.last_kernel_ms
    iu = np.frombuffer(d_info.cpu().numpy().tobytes(), dtype=INFO_DTYPE).copy()
t_c, t_e = ev[0].elapsed_time(ev[1]), ev[2].elapsed_time(ev[3])
gb = (b.qp.nbytes + nb * cl.qp_stride * 8) / 1e9
print(f"batch {nb}: condense {t_c:.2f} ms ({gb / t_c * 1e3:.0f} GB/s of records in+out), expand {t_e:.2f} ms, solve condensed {t_sc:.1f} ms "
      f"(iters {ic['iter'].mean():.2f}, ok {np.mean(ic['status'] == 0):.3f}), solve uncondensed {t_su:.1f} ms (iters {iu['iter'].mean():.2f}, ok {np.mean(iu['status'] == 0):.3f})")
ok = (ic["status"] == 0) & (iu["status"] == 0)
du = np.max(np.abs(b.layout.u_traj(d_s.cpu().numpy()) - b.layout.u_traj(d_s_un.cpu().numpy()))[ok])
print(f"max |u_condensed_path - u_uncondensed_path| over converged instances: {du:.2e} (two IPM trajectories, default tolerances)")
print(f"total condensed path {t_c + t_sc + t_e:.1f} ms -> {nb / (t_c + t_sc + t_e) * 1e3:.0f} QP/s; uncondensed {nb / t_su * 1e3:.0f} QP/s")
