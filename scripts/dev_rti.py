"""Timing of the SQP-RTI split on the device (preparation: condense_lhs; feedback: condense_rhs + solve + expand) for BASELINE
config 3 (pendulum-sized nx=4 nu=1 N=20, batch 16384, N2=5) and config 4 (quadrotor-sized nx=12 nu=4 N=50 -> N2=10, batch 8192);
writes gpurun_out/<tag>_rti.json:  python scripts/dev_rti.py [tag]"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from acados_b200 import problems as P
from acados_b200.binding import CuipmCondenser, CuipmSolver, default_opts, INFO_DTYPE

tag = sys.argv[1] if len(sys.argv) > 1 else "r02"
out = {}
for name, nb, n2 in (("c3", 16384, 5), ("c4", 8192, 10)):
    b = P.named_config(name, nb)
    o = default_opts()
    dc = CuipmCondenser(b.shape, n2)
    cl = dc.condensed_layout
    d_qp = torch.from_numpy(b.qp).cuda()
    # feedback-phase records: same matrices, other vectors (new x0 folded into b / rq of stage 0, new gradients)
    q2 = b.qp.copy()
    rng = np.random.default_rng(1)
    for k in range(b.shape.N + 1):
        for f in ("b", "rq"):
            v = b.layout.view(q2, f, k)
            v += 0.01 * rng.standard_normal(v.shape)
    d_qp2 = torch.from_numpy(q2).cuda()
    d_c = torch.empty((nb, cl.qp_stride), dtype=torch.float64, device="cuda")
    d_sc = torch.zeros((nb, cl.sol_stride), dtype=torch.float64, device="cuda")
    d_s = torch.zeros((nb, b.layout.sol_stride), dtype=torch.float64, device="cuda")
    d_su = torch.zeros((nb, b.layout.sol_stride), dtype=torch.float64, device="cuda")
    d_info = torch.zeros(nb * INFO_DTYPE.itemsize, dtype=torch.uint8, device="cuda")
    su, sc = CuipmSolver(b.shape, nb), CuipmSolver(dc.condensed_shape, nb)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(6)]
    for rep in range(3):
        torch.cuda.synchronize()
        ev[0].record(); dc.condense_lhs(nb, d_qp.data_ptr(), d_c.data_ptr()); ev[1].record()
        torch.cuda.synchronize()
        ev[2].record(); dc.condense_rhs(nb, d_qp2.data_ptr(), d_c.data_ptr()); ev[3].record()
        torch.cuda.synchronize()
        sc.solve_device(nb, d_c.data_ptr(), d_sc.data_ptr(), d_info.data_ptr(), o, sync=True); t_sc = sc.last_kernel_ms
        ic = np.frombuffer(d_info.cpu().numpy().tobytes(), dtype=INFO_DTYPE).copy()
        ev[4].record(); dc.expand(nb, d_qp2.data_ptr(), d_sc.data_ptr(), d_s.data_ptr()); ev[5].record()
        torch.cuda.synchronize()
        su.solve_device(nb, d_qp2.data_ptr(), d_su.data_ptr(), d_info.data_ptr(), o, sync=True); t_su = su.last_kernel_ms
        iu = np.frombuffer(d_info.cpu().numpy().tobytes(), dtype=INFO_DTYPE).copy()
    t_l, t_r, t_e = ev[0].elapsed_time(ev[1]), ev[2].elapsed_time(ev[3]), ev[4].elapsed_time(ev[5])
    ok = (ic["status"] == 0) & (iu["status"] == 0)
    du = float(np.max(np.abs(b.layout.u_traj(d_s.cpu().numpy()) - b.layout.u_traj(d_su.cpu().numpy()))[ok]))
    gb = (b.qp.nbytes + nb * cl.qp_stride * 8) / 1e9
    rec = {"shape": f"nx={b.shape.nx[1]} nu={b.shape.nu[1]} N={b.shape.N} -> N2={n2}", "batch": nb,
           "preparation_condense_lhs_ms": t_l, "lhs_records_gbs": gb / t_l * 1e3,
           "feedback_condense_rhs_ms": t_r, "feedback_solve_condensed_ms": t_sc, "feedback_expand_ms": t_e,
           "feedback_total_ms": t_r + t_sc + t_e, "feedback_qp_per_s": nb / (t_r + t_sc + t_e) * 1e3,
           "one_pass_condensed_path_qp_per_s": nb / (t_l + t_sc + t_e) * 1e3,
           "uncondensed_solve_ms": t_su, "uncondensed_qp_per_s": nb / t_su * 1e3,
           "iters_condensed": float(ic["iter"].mean()), "iters_uncondensed": float(iu["iter"].mean()),
           "converged_frac": float(ok.mean()), "max_du_condensed_vs_uncondensed_path": du,
           "condensed_solver_kernel": "throughput" if sc.last_launch_count > 1 else "generic"}
    out[name] = rec
    print(name, json.dumps(rec), flush=True)
    su.close(); sc.close(); dc.close()
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open(f"gpurun_out/{tag}_rti.json", "w"), indent=1)
