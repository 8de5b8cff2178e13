"""The reference's own golden QP vectors (examples/acados_python/tests/qp_test/last_qp_*.json + sqp_sol_*.json,
asserted there with atol 1e-5 on lam and pi for PARTIAL_CONDENSING_HPIPM) as converted by
tests/golden/make_reference_json_golden.py.  CPU: the oracle; GPU: the CUDA path through the C ABI."""
import os

import numpy as np
import pytest

from acados_b200 import problems as P
from acados_b200.binding import default_opts

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
NAMES = ["nonuniform_pendulum", "one_sided_test"]


def load(name):
    g = np.load(os.path.join(GOLD, f"refjson_{name}.npz"))
    N = int(g["N"]); nb = g["nb"].tolist(); flat = g["idxb"].tolist(); idxb = []; o = 0
    for n in nb:
        idxb.append(flat[o:o + n]); o += n
    shape = P.Shape(N, g["nx"].tolist(), g["nu"].tolist(), nb, [0] * (N + 1), [0] * (N + 1), idxb, [[-1] * n for n in nb])
    return g, P.Batch(shape, P.Layout(shape), np.ascontiguousarray(g["qp"]), name)


def check(g, b, sol, info):
    lay, N = b.layout, b.shape.N
    lam = np.concatenate([lay.view(sol, "lam", k)[0] for k in range(N + 1)])
    pi = np.concatenate([lay.view(sol, "pi", k)[0] for k in range(N)])
    assert info["status"][0] == 0
    assert np.allclose(lam, g["exp_lam"], atol=1e-5) and np.allclose(pi, g["exp_pi"], atol=1e-5)   # the reference's own bar
    assert info["iter"][0] == g["ref_iter"][0]
    assert np.max(np.abs(lay.u_traj(sol) - lay.u_traj(g["ref_sol"]))) <= 1e-10                    # parity with HPIPM itself


@pytest.mark.parametrize("name", NAMES)
def test_oracle_on_reference_json_golden(built, name):
    from oracle import oracle_binding as ob
    g, b = load(name)
    sol, info = ob.oracle_solve(b, default_opts(iter_max=500))
    check(g, b, sol, info)


@pytest.mark.gpu
@pytest.mark.parametrize("name", NAMES)
def test_cuda_on_reference_json_golden(built, name):
    from acados_b200.binding import CuipmSolver
    g, b = load(name)
    s = CuipmSolver(b.shape, 1)
    sol, info = s.solve(b.qp, default_opts(iter_max=500))
    s.close()
    check(g, b, sol, info)
