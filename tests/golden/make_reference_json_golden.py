"""Turns the reference's own golden QP fixtures (examples/acados_python/tests/qp_test/last_qp_*.json with the expected
solutions sqp_sol_*.json, used by tests/qp_test/test_ocpqp_solver.py with atol 1e-5 on lam and pi) into cuipm records.

The JSON holds the QP before condensing: stage 0 carries x0 as equality boxes (idxe).  The reference eliminates them
before calling the QP solver (d_ocp_qp_reduce_eq_dof, external/hpipm/ocp_qp/x_ocp_qp_red.c:278); this script does the
same elimination in numpy (b_0 += A_0 x0, r_0 += S_0 x0, stage 0 keeps only u), checks with the compiled reference
(oracle/_ref) that the reduced QP reproduces the expected lam / pi / u to 1e-5, and stores qp record + expectations.
Run here (needs /root/reference); the .npz files are committed.
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from acados_b200 import problems as P  # noqa: E402
from acados_b200.binding import default_opts  # noqa: E402
from oracle import oracle_binding as ob  # noqa: E402

SRC = "/root/reference/examples/acados_python/tests/qp_test"
here = os.path.dirname(os.path.abspath(__file__))


def find(d, key, k):
    for cand in (f"{key}_{k}", f"{key}_{k:02d}", f"{key}_{k:03d}"):
        if cand in d:
            return cand
    return None


def arr(d, key, k, dtype=float):
    c = find(d, key, k)
    return np.array(d[c], dtype=dtype) if c else np.zeros((0,))


def convert(name):
    d = json.load(open(os.path.join(SRC, f"last_qp_{name}.json")))
    e = json.load(open(os.path.join(SRC, f"sqp_sol_{name}.json")))
    N = max(int(k.split("_")[1]) for k in d if k.startswith("A_")) + 1
    nx = [arr(d, "Q", k).shape[0] for k in range(N + 1)]
    nu = [arr(d, "R", k).shape[0] if find(d, "R", k) and k < N else 0 for k in range(N + 1)]
    x0 = arr(d, "lbx", 0).ravel()
    assert np.allclose(x0, arr(d, "ubx", 0).ravel()) and len(x0) == nx[0], "stage-0 state box must pin x0"
    nxs = [0] + nx[1:]
    idxb, nb, lbs, ubs, lms, ums = [], [], [], [], [], []
    for k in range(N + 1):
        lbu, ubu = arr(d, "lbu", k).ravel(), arr(d, "ubu", k).ravel()
        lbx, ubx = (arr(d, "lbx", k).ravel(), arr(d, "ubx", k).ravel()) if k > 0 else (np.zeros(0), np.zeros(0))
        ib = arr(d, "idxb", k, int).ravel().tolist()
        nbu = len(lbu)
        ib = ib[:nbu] + (ib[nbu:nbu + len(lbx)] if k > 0 else [])
        idxb.append(ib); nb.append(len(ib))
        lbs.append(np.concatenate([lbu, lbx])); ubs.append(np.concatenate([ubu, ubx]))
        mk = lambda key, n: (arr(d, key, k).ravel() if find(d, key, k) else np.ones(n))
        lms.append(np.concatenate([mk("lbu_mask", nbu), mk("lbx_mask", len(lbx)) if k > 0 else np.zeros(0)]))
        ums.append(np.concatenate([mk("ubu_mask", nbu), mk("ubx_mask", len(lbx)) if k > 0 else np.zeros(0)]))
    shape = P.Shape(N, nxs, nu, nb, [0] * (N + 1), [0] * (N + 1), idxb, [[-1] * n for n in nb])
    lay = P.Layout(shape)
    qp = lay.new_qp(1)
    for k in range(N + 1):
        Q, q = arr(d, "Q", k), arr(d, "q", k).ravel()
        R = arr(d, "R", k) if nu[k] else np.zeros((0, 0))
        S = arr(d, "S", k) if nu[k] else np.zeros((0, nx[k]))       # nu x nx
        r = arr(d, "r", k).ravel() if nu[k] else np.zeros(0)
        if k == 0:
            P._set_cost(lay, qp, 0, R=R[None], r=(r + S @ x0)[None])
            A, B, b = arr(d, "A", 0), arr(d, "B", 0), arr(d, "b", 0).ravel()
            P._set_dyn(lay, qp, 0, B=B[None], b=(b + A @ x0)[None])
        else:
            P._set_cost(lay, qp, k, R=R[None] if nu[k] else None, S=S.T[None] if nu[k] else None, Q=Q[None],
                        r=r[None] if nu[k] else None, q=q[None])
            if k < N:
                P._set_dyn(lay, qp, k, A=arr(d, "A", k)[None], B=arr(d, "B", k)[None], b=arr(d, "b", k).ravel()[None])
        if nb[k]:
            P._set_box(lay, qp, k, lbs[k][None], ubs[k][None], lb_mask=lms[k][None], ub_mask=ums[k][None])
    b = P.Batch(shape, lay, qp, name)
    # expectations from the reference's stored SQP solution (reduced to the variables that survive the elimination)
    exp_u = np.concatenate([arr(e, "u", k).ravel() for k in range(N)])
    exp_pi = np.concatenate([arr(e, "pi", k).ravel() for k in range(N)])
    exp_lam = []
    for k in range(N + 1):
        lam = arr(e, "lam", k).ravel()
        if k == 0:   # full stage 0 has nbu + nx boxes: keep the input bounds only
            nbf = len(lam) // 2
            nbu = len(arr(d, "lbu", 0).ravel())
            lam = np.concatenate([lam[:nbu], lam[nbf:nbf + nbu]])
        exp_lam.append(lam)
    exp_lam = np.concatenate(exp_lam)
    o = default_opts(iter_max=500)
    sol, info, _ = ob.ref_solve(b, o, nthreads=1)
    got_lam = np.concatenate([lay.view(sol, "lam", k)[0] for k in range(N + 1)])
    got_pi = np.concatenate([lay.view(sol, "pi", k)[0] for k in range(N)])
    print(name, "N", N, "status", info["status"], "iter", info["iter"], "|u-u*|", np.max(np.abs(lay.u_traj(sol)[0] - exp_u)),
          "|pi-pi*|", np.max(np.abs(got_pi - exp_pi)), "|lam-lam*|", np.max(np.abs(got_lam - exp_lam)))
    assert info["status"][0] == 0 and np.allclose(got_lam, exp_lam, atol=1e-5) and np.allclose(got_pi, exp_pi, atol=1e-5)
    np.savez_compressed(os.path.join(here, f"refjson_{name}.npz"), N=N, nx=nxs, nu=nu, nb=nb, idxb=np.concatenate([np.array(i, int) for i in idxb]),
                        qp=qp, exp_u=exp_u, exp_pi=exp_pi, exp_lam=exp_lam, ref_sol=sol, ref_iter=info["iter"])


for n in ("nonuniform_pendulum", "one_sided_test"):
    convert(n)
