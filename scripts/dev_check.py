"""Development check on a GPU box: CUDA path vs CPU oracle (and the reference, when oracle/_ref is present)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from acados_b200 import problems as P
from acados_b200.binding import CuipmSolver, default_opts
from oracle import oracle_binding as O

names = sys.argv[1:] or ["c1", "c3", "c2", "c4", "c5"]
nb = {"c1": 8, "c2": 16, "c3": 32, "c4": 8, "c5": 4}
for name in names:
    b = P.named_config(name, nb.get(name, 8))
    o = default_opts()
    s = CuipmSolver(b.shape, b.nbatch)
    for W in (1, 2, 4):
        s.set_tuning("warps", W)
        t0 = time.time()
        sol, info, stat = s.solve(b.qp, o, want_stat=True)
        t1 = time.time()
        osol, oinfo, ostat = O.oracle_solve(b, o, want_stat=True)
        du = np.max(np.abs(b.layout.u_traj(sol) - b.layout.u_traj(osol)))
        ds = np.max(np.abs(sol - osol))
        print(f"{name} W={W} iters {info['iter'][:8]} oracle {oinfo['iter'][:8]} status {info['status'][:8]} du {du:.2e} dsol {ds:.2e} "
              f"res {info['res_max'].max(0)} kernel_ms {s.last_kernel_ms:.3f} wall {t1-t0:.3f}", flush=True)
        if not np.array_equal(info['iter'], oinfo['iter']) or not du < 1e-9:
            q = int(np.argmax(np.abs(b.layout.u_traj(sol) - b.layout.u_traj(osol)).max(1)))
            print("  first rows of stat (gpu / oracle) for qp", q)
            for r in range(min(4, stat.shape[1])):
                print("   g", np.array2string(stat[q, r, :13], precision=4))
                print("   o", np.array2string(ostat[q, r, :13], precision=4))
    s.close()
