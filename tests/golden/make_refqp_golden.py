"""Converts the reference's golden QP fixtures (examples/acados_python/tests/qp_test/last_qp_*.json, expected multipliers
in sqp_sol_*.json; asserted by the reference's tests/qp_test/test_ocpqp_solver.py with atol 1e-5 on lam and pi) into one
JSON per case for tests/test_ocp_qp_mirror.py: the QP dictionary UNREDUCED (x0 still a stage-0 equality, the form a user
hands to AcadosOcpQp.from_json), the expected lam / pi, and -- from the compiled reference (oracle/_ref) run on the
x0-eliminated records -- the iteration count and input trajectory of HPIPM itself.
Run here (needs /root/reference); the outputs are committed."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from acados_b200.binding import default_opts  # noqa: E402
from acados_b200.ocp_qp import OcpQp, PackedBatch  # noqa: E402
from acados_b200.problems import Batch  # noqa: E402
from oracle import oracle_binding as ob  # noqa: E402

SRC = "/root/reference/examples/acados_python/tests/qp_test"
here = os.path.dirname(os.path.abspath(__file__))

for name in ("nonuniform_pendulum", "one_sided_test"):
    d = json.load(open(os.path.join(SRC, f"last_qp_{name}.json")))
    e = json.load(open(os.path.join(SRC, f"sqp_sol_{name}.json")))
    qp = OcpQp.from_json(json_data=dict(d))
    p = PackedBatch([qp])
    b = Batch(p.shape, p.layout, p.qp, name)
    sol, info, _ = ob.ref_solve(b, default_opts(iter_max=500), nthreads=1)
    assert info["status"][0] == 0
    w = len(str(qp.N + 1))
    out = {"qp": d, "exp_lam": [e.get(f"lam_{k:0{w}d}", []) for k in range(qp.N + 1)], "exp_pi": [e[f"pi_{k:0{w}d}"] for k in range(qp.N)],
           "ref_iter": int(info["iter"][0]), "ref_u": p.layout.u_traj(sol)[0].tolist()}
    json.dump(out, open(os.path.join(here, f"refqp_{name}.json"), "w"))
    print(name, "N", qp.N, "iter", out["ref_iter"], "nx", qp.dims.nx.tolist()[:3], "nbxe0", int(qp.dims.nbxe[0]))
