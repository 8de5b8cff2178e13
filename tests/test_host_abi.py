"""CPU tests of the host side of the C ABI: layout, options, exported symbols (no compute calls, no GPU)."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

from acados_b200 import problems as P
from acados_b200.binding import (CuipmOpts, c_layout_as_dict, default_opts, load_library)

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _shapes():
    yield P.mass_spring(1).shape
    yield P.chain_mass(1, N=6).shape
    yield P.random_shape(5, 4, 2, nbx=2, ng=2, ns=2)
    yield P.random_shape(3, 3, 1, nbx=3, ng=1, ns=3, x0_eliminated=False, terminal_nu=1)


def test_every_declared_symbol_is_exported(built):
    hdr = open(os.path.join(ROOT, "include", "cuipm.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    names = set(re.findall(r"\b(cuipm_[a-z_]+)\s*\(", hdr))
    assert len(names) >= 18
    out = subprocess.run(["nm", "-D", "--defined-only", os.path.join(ROOT, "acados_b200", "csrc", "libcuipm.so")],
                         stdout=subprocess.PIPE, text=True, check=True).stdout
    exported = set(re.findall(r" T (cuipm_[a-z_]+)", out))
    missing = names - exported
    assert not missing, f"declared in include/cuipm.h but not exported: {sorted(missing)}"
    lib = load_library()
    for n in names:
        assert hasattr(lib, n)


def test_library_does_not_depend_on_oracle(built):
    out = subprocess.run(["ldd", os.path.join(ROOT, "acados_b200", "csrc", "libcuipm.so")], stdout=subprocess.PIPE, text=True).stdout
    assert "oracle" not in out and "acados_ref" not in out and "hpipm" not in out and "blasfeo" not in out


@pytest.mark.parametrize("shape", list(_shapes()))
def test_layout_three_ways(built, shape):
    """numpy Layout == cuipm_layout_create (product) == oracle_layout_create (independent restatement)."""
    from oracle import oracle_binding as ob
    lib = load_library()
    lay = P.Layout(shape)
    p = lib.cuipm_layout_create(C.byref(shape.as_ctypes()))
    prod = c_layout_as_dict(p, shape.N)
    lib.cuipm_layout_destroy(p)
    orc = ob.oracle_layout(shape)
    assert prod == orc
    assert prod["qp_stride"] == lay.qp_stride and prod["sol_stride"] == lay.sol_stride
    for cname, pname in (("off_BAt", "BAt"), ("off_RSQ", "RSQ"), ("off_DCt", "DCt"), ("off_b", "b"), ("off_rq", "rq"),
                         ("off_d", "d"), ("off_dmask", "dmask"), ("off_Z", "Z"), ("off_z", "z"), ("off_ux", "ux"),
                         ("off_pi", "pi"), ("off_lam", "lam"), ("off_t", "t")):
        assert prod[cname] == lay.off[pname], cname
    assert all(o % 2 == 0 for v in lay.off.values() for o in v)      # 16-byte aligned sub-arrays
    assert prod["qp_stage"] == lay.qp_stage and prod["sol_stage"] == lay.sol_stage


def test_option_defaults_match_reference_values(built):
    """BALANCE defaults (external/hpipm/ocp_qp/x_ocp_qp_ipm.c:145-179) + acados overrides (acados/ocp_qp/ocp_qp_hpipm.c:101-113)."""
    o = default_opts("BALANCE", acados=False)
    assert (o.mu0, o.alpha_min, o.iter_max, o.itref_corr_max, o.lq_fact, o.split_step) == (10.0, 1e-12, 30, 2, 1, 0)
    assert (o.res_g_max, o.res_b_max, o.res_d_max, o.res_m_max, o.reg_prim) == (1e-6, 1e-8, 1e-8, 1e-8, 1e-15)
    assert (o.t_lam_min, o.t0_init, o.var_init_scheme, o.pred_corr, o.cond_pred_corr) == (2, 2, 0, 1, 1)
    a = default_opts("BALANCE", acados=True)
    assert (a.iter_max, a.stat_max, a.alpha_min, a.mu0, a.var_init_scheme) == (50, 50, 1e-8, 1.0, 1)
    r = default_opts("ROBUST", acados=False)
    assert (r.mu0, r.iter_max, r.itref_corr_max, r.lq_fact) == (100.0, 100, 4, 2)
    s = default_opts("SPEED", acados=False)
    assert (s.iter_max, s.split_step, s.lq_fact) == (15, 1, 0)


def test_opts_set_get_by_reference_field_names(built):
    lib = load_library()
    o = default_opts()
    for name, val, attr in (("iter_max", 77, "iter_max"), ("warm_start", 1, "warm_start")):
        assert lib.cuipm_opts_set(C.byref(o), name.encode(), C.byref(C.c_int(val))) == 0
        assert getattr(o, attr) == val
    assert o.stat_max >= o.iter_max
    for name, val, attr in (("tol_stat", 1e-9, "res_g_max"), ("tol_eq", 1e-10, "res_b_max"), ("tol_ineq", 1e-11, "res_d_max"),
                            ("tol_comp", 1e-12, "res_m_max"), ("mu0", 3.0, "mu0")):
        assert lib.cuipm_opts_set(C.byref(o), name.encode(), C.byref(C.c_double(val))) == 0
        assert getattr(o, attr) == val
        got = C.c_double()
        assert lib.cuipm_opts_get(C.byref(o), name.encode(), C.byref(got)) == 0 and got.value == val
    assert lib.cuipm_opts_set(C.byref(o), b"hpipm_mode", C.c_char_p(b"SPEED")) == 0 and o.mode == 1 and o.iter_max == 50
    assert lib.cuipm_opts_set(C.byref(o), b"no_such_field", C.byref(C.c_int(1))) != 0
    assert lib.cuipm_opts_set(C.byref(o), b"hpipm_mode", C.c_char_p(b"NOPE")) != 0


def test_create_without_gpu_fails_loudly(built):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from acados_b200.binding import CuipmSolver
    with pytest.raises(RuntimeError, match="no CUDA device|CUDA"):
        CuipmSolver(P.mass_spring(1).shape, 1)


def test_reducer_and_condenser_without_gpu_fail_loudly(built):
    """The device-side elimination / condensing objects need a CUDA device too: no host fallback behind the C ABI."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from acados_b200.binding import CuipmCondenser, CuipmReducer
    sh = P.chain_mass(1, N=6).shape
    with pytest.raises(RuntimeError, match="CUDA"):
        CuipmCondenser(sh, 3)
    full = P.random_shape(4, 3, 2, nbx=3, x0_eliminated=False)
    with pytest.raises(RuntimeError, match="CUDA"):
        CuipmReducer(full, [full.nu[0]])


def test_problem_generators_are_deterministic_and_well_formed():
    a, b = P.chain_mass(3, N=5, seed=7), P.chain_mass(3, N=5, seed=7)
    assert np.array_equal(a.qp, b.qp)
    c = P.chain_mass(3, N=5, seed=8)
    assert not np.array_equal(a.qp, c.qp)
    lay = a.layout
    for k in range(a.shape.N + 1):
        H = lay.view(a.qp, "RSQ", k)
        assert np.allclose(H, np.swapaxes(H, 1, 2))
        assert (np.linalg.eigvalsh(H) > 0).all()
    ms = P.mass_spring(1)
    assert ms.shape.nx == [0] + [8] * 15 and ms.shape.nu == [3] * 15 + [0] and ms.shape.nb[1] == 11
