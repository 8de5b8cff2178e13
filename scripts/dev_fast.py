"""Development check of the throughput kernel on a GPU box: parity against the oracle and the generic kernel, timing."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from acados_b200 import problems
from acados_b200.binding import CuipmSolver, default_opts, INFO_DTYPE
from oracle import oracle_binding as ob

def parity(b, name, **ov):
    o = default_opts(**ov)
    s = CuipmSolver(b.shape, b.nbatch)
    sol, info, stat = s.solve(b.qp, o, want_stat=True)
    nl = s.last_launch_count
    s.set_tuning("fast", 0)
    gsol, ginfo, gstat = s.solve(b.qp, o, want_stat=True)
    osol, oinfo = ob.oracle_solve(b, o, nthreads=8)
    lay = b.layout
    du = np.max(np.abs(lay.u_traj(sol) - lay.u_traj(osol)))
    dug = np.max(np.abs(lay.u_traj(sol) - lay.u_traj(gsol)))
    print(f"{name}: launches {nl} iter_eq_oracle {(info['iter']==oinfo['iter']).mean():.3f} status_eq {(info['status']==oinfo['status']).mean():.3f} "
          f"du_oracle {du:.2e} du_generic {dug:.2e} dsol_generic {np.max(np.abs(sol-gsol)):.2e} iters {info['iter'][:8]} status {np.bincount(info['status'])}", flush=True)
    s.close()

def timing(b, name, steps=3):
    o = default_opts()
    nb = b.nbatch
    s = CuipmSolver(b.shape, nb)
    d_qp = torch.from_numpy(b.qp).cuda()
    d_sol = torch.zeros((nb, b.layout.sol_stride), dtype=torch.float64, device="cuda")
    d_info = torch.zeros(nb * INFO_DTYPE.itemsize, dtype=torch.uint8, device="cuda")
    for fast, rr in ((1, 1), (1, 0), (0, 0)):
        s.set_tuning("fast", fast)
        s.set_tuning("rr", rr)
        s.solve_device(nb, d_qp.data_ptr(), d_sol.data_ptr(), d_info.data_ptr(), o, sync=True)
        ms = []
        for _ in range(steps):
            s.solve_device(nb, d_qp.data_ptr(), d_sol.data_ptr(), d_info.data_ptr(), o, sync=True)
            ms.append(s.last_kernel_ms)
        info = np.frombuffer(d_info.cpu().numpy().tobytes(), dtype=INFO_DTYPE)
        print(f"{name} batch {nb} fast={fast} rr={rr} launches {s.last_launch_count}: {min(ms):.2f} ms -> {nb/min(ms)*1e3:.0f} QP/s  iters {info['iter'].mean():.2f} status {np.bincount(info['status'])}", flush=True)
    s.close()

if __name__ == "__main__":
    what = sys.argv[1] if len(sys.argv) > 1 else "all"
    if what in ("all", "parity"):
        parity(problems.chain_mass(64, N=40, seed=5), "c2")
        parity(problems.mass_spring(64, seed=1, x0_scale=0.5), "c1")
        parity(problems.named_config("c3", 64), "c3")
        parity(problems.named_config("c4", 64), "c4")
        sh = problems.random_shape(12, 8, 3, nbx=4, ns=2)
        parity(problems.random_qp(sh, 64, seed=5, mask_frac=0.3), "rand_soft_mask")
        parity(problems.chain_mass(64, N=40, seed=5), "c2 lq0", lq_fact=0)
    if what == "ncu":
        # one launch per kernel for a profiler: python scripts/dev_fast.py ncu <config> <batch>
        cfg, nb = sys.argv[2], int(sys.argv[3])
        b = problems.chain_mass(nb, N=40, seed=1234) if cfg == "c2" else problems.named_config(cfg, nb)
        o = default_opts()
        s = CuipmSolver(b.shape, nb)
        d_qp = torch.from_numpy(b.qp).cuda()
        d_sol = torch.zeros((nb, b.layout.sol_stride), dtype=torch.float64, device="cuda")
        d_info = torch.zeros(nb * INFO_DTYPE.itemsize, dtype=torch.uint8, device="cuda")
        s.solve_device(nb, d_qp.data_ptr(), d_sol.data_ptr(), d_info.data_ptr(), o, sync=True)
        print("kernel ms", s.last_kernel_ms)
        s.close()
    if what == "prof":
        # cycles per sweep (library built with CUIPM_DEFS=-DFK_PROFILE): python scripts/dev_fast.py prof <config> <batch>
        import ctypes as C
        cfg, nb = sys.argv[2], int(sys.argv[3])
        b = problems.chain_mass(nb, N=40, seed=1234) if cfg == "c2" else problems.named_config(cfg, nb)
        o = default_opts()
        s = CuipmSolver(b.shape, nb)
        d_qp = torch.from_numpy(b.qp).cuda()
        d_sol = torch.zeros((nb, b.layout.sol_stride), dtype=torch.float64, device="cuda")
        d_info = torch.zeros(nb * INFO_DTYPE.itemsize, dtype=torch.uint8, device="cuda")
        s.solve_device(nb, d_qp.data_ptr(), d_sol.data_ptr(), d_info.data_ptr(), o, sync=True)
        out = (C.c_ulonglong * 32)()
        s.lib.cuipm_fast_prof_read(out, 1)
        s.solve_device(nb, d_qp.data_ptr(), d_sol.data_ptr(), d_info.data_ptr(), o, sync=True)
        s.lib.cuipm_fast_prof_read(out, 1)
        v = [float(x) for x in out]
        names = ["res", "fact", "fwd", "solve", "mu_aff", "wait_vec", "wait_mat", "total", "warps"]
        print(f"kernel {s.last_main_kernel_ms:.2f} ms; cycles per warp: total {v[7]/v[8]:.3g}")
        for i in range(7):
            print(f"  {names[i]:9s} {100*v[i]/v[7]:6.2f} % of the warp time")
        for i, nm in ((10, "fact: prologue + waits"), (11, "fact: gradient (alb, Pb)"), (12, "fact: TRMM"), (13, "fact: SYRK + update + H"), (14, "fact: panels"), (16, "fwd: issue + waits"), (17, "fwd: u"), (18, "fwd: H v"),
                      (19, "fwd: x+"), (20, "fwd: constraint step"), (21, "fwd: pi"), (22, "fwd: residual rows")):
            print(f"  {nm:26s} {100*v[i]/v[7]:6.2f} %")
        s.close()
    if what == "time3":
        timing(problems.chain_mass(4096, N=40, seed=1234), "c2")
        timing(problems.named_config("c3", 16384), "c3")
    if what == "rr":
        # iteration-sliced scheduling on / off: bit identity and timing
        for nbq in (4096, 8192, 2048):
            b = problems.chain_mass(nbq, N=40, seed=1234)
            o = default_opts()
            s = CuipmSolver(b.shape, nbq)
            d_qp = torch.from_numpy(b.qp).cuda()
            outs = {}
            for rr in (1, 0, 1, 0):
                s.set_tuning("rr", rr)
                d_sol = torch.zeros((nbq, b.layout.sol_stride), dtype=torch.float64, device="cuda")
                d_info = torch.zeros(nbq * INFO_DTYPE.itemsize, dtype=torch.uint8, device="cuda")
                ms = []
                for _ in range(3):
                    s.solve_device(nbq, d_qp.data_ptr(), d_sol.data_ptr(), d_info.data_ptr(), o, sync=True)
                    ms.append(s.last_kernel_ms)
                outs[rr] = (d_sol.cpu().numpy(), d_info.cpu().numpy().copy())
                print(f"c2 batch {nbq} rr={rr}: {min(ms):.2f} ms -> {nbq/min(ms)*1e3:.0f} QP/s launches {s.last_launch_count}", flush=True)
            print("   bit-identical:", np.array_equal(outs[0][0], outs[1][0]), np.array_equal(outs[0][1], outs[1][1]), flush=True)
            s.close()
    if what == "time5":
        timing(problems.named_config("c5", 1024), "c5")
    if what == "par5":
        parity(problems.named_config("c5", 64), "c5")
    if what == "time2":
        timing(problems.chain_mass(4096, N=40, seed=1234), "c2")
    if what in ("all", "time"):
        timing(problems.chain_mass(4096, N=40, seed=1234), "c2")
        timing(problems.chain_mass(8192, N=40, seed=1234), "c2")
        timing(problems.named_config("c3", 16384), "c3")
        timing(problems.named_config("c4", 8192), "c4")
        timing(problems.mass_spring(16384, seed=1, x0_scale=0.5), "c1")
