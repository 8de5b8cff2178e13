"""Per-pass cycle split (needs a CUIPM_PROFILE=1 build): python scripts/dev_prof.py c2 [nbatch]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from acados_b200 import problems as P
from acados_b200.binding import CuipmSolver, default_opts
name = sys.argv[1]; nb = int(sys.argv[2]) if len(sys.argv) > 2 else 1
b = P.named_config(name, nb); o = default_opts()
s = CuipmSolver(b.shape, nb)
for W in (1,):
    s.set_tuning("warps", W)
    sol, info, stat = s.solve(b.qp, o, want_stat=True)
    sol, info, stat = s.solve(b.qp, o, want_stat=True)
    pr = stat[:, o.stat_max, :16].mean(0)
    names = ["res", "res_lin", "fact_bwd", "forward", "solve_bwd", "vector"]
    tot = pr[:6].sum()
    print(f"{name} nb={nb} W={W} kernel {s.last_kernel_ms:.2f} ms iters {info['iter'].mean():.1f}; cycles/QP {tot:.3e}")
    for n_, c in zip(names, pr[:6]):
        print(f"   {n_:10s} {c:12.0f} cycles  {100*c/tot:5.1f}%   per iter {c/info['iter'].mean():10.0f}")
