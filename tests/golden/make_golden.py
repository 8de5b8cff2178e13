"""Generates the golden vectors in this directory FROM THE REFERENCE ITSELF (oracle/_ref, i.e. unmodified
HPIPM+BLASFEO behind acados' qp_solver vtable, built from /root/reference by oracle/Makefile).  Run here, where the
reference exists; the .npz files are committed and travel to the GPU box, where /root/reference does not exist.
  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from acados_b200.binding import default_opts  # noqa: E402
from oracle import oracle_binding as ob  # noqa: E402
from test_oracle_vs_reference import CASES  # noqa: E402

assert ob.have_ref(), "build oracle/_ref first: make -C oracle"
here = os.path.dirname(os.path.abspath(__file__))
for name, make in CASES.items():
    b = make()
    sol, info, _ = ob.ref_solve(b, default_opts(), nthreads=1)
    np.savez_compressed(os.path.join(here, name + ".npz"), case=name, qp_head=np.asarray(b.qp[:, :64]), sol=sol,
                        u=b.layout.u_traj(sol), iter=info["iter"], status=info["status"], res_max=info["res_max"])
    print(name, "iters", info["iter"].tolist(), "status", info["status"].tolist())
