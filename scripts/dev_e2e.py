"""End-to-end probe of cuipm_solve_host (pinned host buffers) for different chunk counts: python scripts/dev_e2e.py c2 4096"""
import sys, os, time, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from acados_b200 import problems as P
from acados_b200.binding import CuipmSolver, default_opts, INFO_DTYPE
name = sys.argv[1]; nb = int(sys.argv[2])
b = P.named_config(name, nb)
o = default_opts()
s = CuipmSolver(b.shape, nb)
h_qp = torch.from_numpy(b.qp).pin_memory()
h_sol = torch.zeros((nb, b.layout.sol_stride), dtype=torch.float64).pin_memory()
h_info = torch.zeros(nb * INFO_DTYPE.itemsize, dtype=torch.uint8).pin_memory()
for pipe in (1, 2, 4, 6, 8):
    s.set_tuning("pipe", pipe)
    ts = []
    for rep in range(4):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        rc = s.lib.cuipm_solve_host(s.handle, nb, h_qp.data_ptr(), h_sol.data_ptr(), h_info.data_ptr(), None, C.byref(o))
        assert rc == 0
        ts.append(time.perf_counter() - t0)
    print(f"pipe={pipe}: {min(ts[1:])*1e3:.1f} ms -> {nb/min(ts[1:]):.0f} QP/s (device span {s.last_kernel_ms:.1f} ms)", flush=True)
