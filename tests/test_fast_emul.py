"""The body of the throughput kernel (acados_b200/csrc/cuipm_fast_core.h) executed on a host emulation of a warp
(oracle/simt_emul.h, oracle/fast_emul.cpp) against the oracle: same records, iteration counts and status codes equal,
|du| <= 1e-10 -- for every lanes-per-QP mapping compiled, both lane orders of the emulation (missing warp barriers show
up in one of them), box / soft / masked constraints, warm starts, tight tolerances.  QPs the kernel hands back to the
generic kernel (cold paths: LQ refactorisation, iterative refinement) are excluded from the comparison and must stay rare.
The GPU suite runs the same body as compiled by nvcc (tests/test_parity_gpu.py)."""
import numpy as np
import pytest

from acados_b200 import problems
from acados_b200.binding import default_opts
from oracle import oracle_binding as ob


def _check(b, g, order=0, max_redo=0, sol0=None, tol_u=1e-10, check_stat=True, sol_rtol=0.0, **ov):
    o = default_opts(**ov)
    sol, info, stat, redo = ob.fast_emul_solve(b, o, g=g, order=order, sol0=sol0, want_stat=True)
    osol, oinfo, ostat = ob.oracle_solve(b, o, sol0=sol0, want_stat=True)
    keep = np.setdiff1d(np.arange(b.nbatch), redo)
    assert len(redo) <= max_redo, redo
    assert (info["status"][redo] == 100).all()
    assert (info["iter"][keep] == oinfo["iter"][keep]).all()
    assert (info["status"][keep] == oinfo["status"][keep]).all()
    lay = b.layout
    du = np.max(np.abs(lay.u_traj(sol) - lay.u_traj(osol))[keep])
    assert du <= tol_u, du
    assert np.max((np.abs(sol - osol) - sol_rtol * np.abs(osol))[keep]) <= max(1e-8, 100 * tol_u)
    for i in (keep[:4] if check_stat else []):
        it = int(oinfo["iter"][i])
        # per-iteration statistics: alpha, mu_aff, sigma, alpha, mu, residual norms, gap, objective (columns 0..12)
        a, r = stat[i, :it + 1, :13], ostat[i, :it + 1, :13]
        assert np.all(np.abs(a - r) <= 1e-4 * np.abs(r) + 1e-7), (i, np.max(np.abs(a - r)))   # residuals at round-off level differ
    return info, redo


@pytest.mark.parametrize("g,order", [(8, 0), (8, 1), (16, 0), (32, 1)])
def test_chain_mass(g, order):
    _check(problems.chain_mass(5, N=10, seed=3), g, order)


def test_chain_mass_full_horizon():
    _check(problems.chain_mass(4, N=40, seed=7), 8)


@pytest.mark.parametrize("g,order", [(4, 0), (8, 1), (16, 0)])
def test_mass_spring(g, order):
    _check(problems.mass_spring(8, seed=1, x0_scale=0.5), g, order)


@pytest.mark.parametrize("g,order", [(2, 0), (4, 1), (8, 0)])
def test_pendulum_sized(g, order):
    _check(problems.named_config("c3", 16), g, order)


@pytest.mark.parametrize("g,order", [(8, 0), (16, 1)])
def test_quadrotor_sized_soft(g, order):
    _check(problems.named_config("c4", 4), g, order, max_redo=1)


def test_legged_sized():
    """nx=48 nu=12 (BASELINE config 5 shape, short horizon): one QP per warp, two row slots per lane, eight column tiles.
    The random family stops ~1e-8 from the solution with multipliers of order 1e2..1e3: held to 1e-9 on u as in the GPU suite."""
    sh = problems.random_shape(4, 48, 12, nbx=12, ns=12)
    b = problems.random_qp(sh, 2, seed=3, umax=0.5, xmax=1.0, x0_scale=1.0)
    _check(b, 32, order=1, tol_u=1e-9, sol_rtol=1e-8, check_stat=False)


@pytest.mark.parametrize("g,order", [(8, 0), (4, 1)])
def test_soft_and_masked(g, order):
    sh = problems.random_shape(8, 8, 3, nbx=4, ns=2)
    _check(problems.random_qp(sh, 8, seed=5, mask_frac=0.3), g, order, max_redo=3)


def test_hard_constraints_and_partial_warp():
    # 3 QPs with 4 QPs per warp: the last group of the warp has no QP of its own
    _check(problems.chain_mass(3, N=8, seed=9, soft=False), 8)


def test_options():
    b = problems.chain_mass(3, N=8, seed=9)
    _check(b, 8, lq_fact=0)
    _check(b, 8, res_g_max=1e-12, res_b_max=1e-12, res_d_max=1e-12, res_m_max=1e-12)
    _check(b, 8, pred_corr=0, iter_max=80, stat_max=80)
    _check(b, 8, cond_pred_corr=0)
    _check(b, 8, t_lam_min=1)
    _check(b, 8, t0_init=0)
    _check(b, 8, iter_max=3, stat_max=3)      # stops at the iteration limit (status 1)


def test_warm_start():
    b = problems.chain_mass(4, N=8, seed=11)
    o = default_opts()
    osol, _ = ob.oracle_solve(b, o)
    for ws in (2, 3):
        # restarting from a converged point puts mu at round-off level: same bounds as tests/test_parity_gpu.py
        _check(b, 8, sol0=osol, warm_start=ws, tol_u=1e-9 if ws == 2 else 1e-7, check_stat=False)


@pytest.mark.parametrize("g,order", [(8, 0), (8, 1), (4, 1)])
def test_iteration_sliced_scheduling_is_bit_identical(g, order):
    """rr_first / rr_loop (a QP bound to a warp for one iteration, the unfinished ones circulating through a ring) against the
    one-warp-per-group schedule: same arithmetic on the same records, hence bit-identical solutions, summaries and statistics
    whatever the order in which the QPs meet in a warp; QPs with different iteration counts, a QP that is handed back."""
    b = problems.chain_mass(7, N=8, seed=4) if g == 8 else problems.mass_spring(11, seed=2, x0_scale=0.5)
    if g == 8:
        for k in range(b.shape.N + 1):
            b.layout.view(b.qp, "dmask", k)[5] = 0.0           # QP 5 has no active constraint: handed back by the first launch
    o = default_opts()
    s0, i0, st0, r0 = ob.fast_emul_solve(b, o, g=g, order=order, want_stat=True)
    s1, i1, st1, r1 = ob.fast_emul_solve(b, o, g=g, order=order, want_stat=True, rr=True)
    assert np.array_equal(r0, r1) and len(set(i0["iter"].tolist())) > 1
    assert np.array_equal(s0, s1) and np.array_equal(st0, st1)
    for f in ("status", "iter", "mu", "obj", "dual_gap", "res_max"):
        assert np.array_equal(i0[f], i1[f]), f


@pytest.mark.parametrize("shape", ["pendulum_g2", "legged_g32"])
def test_iteration_sliced_scheduling_other_mappings(shape):
    """The same with 16 QPs per warp (two lanes per QP) and with one QP per warp (the tensor-core instance)."""
    if shape == "pendulum_g2":
        b, g = problems.named_config("c3", 21), 2
    else:
        b, g = problems.random_qp(problems.random_shape(3, 48, 12, nbx=12, ns=12), 3, seed=3, umax=0.5, xmax=1.0, x0_scale=1.0), 32
    o = default_opts()
    s0, i0, r0 = ob.fast_emul_solve(b, o, g=g)
    s1, i1, r1 = ob.fast_emul_solve(b, o, g=g, rr=True)
    assert np.array_equal(r0, r1) and np.array_equal(s0, s1)
    assert np.array_equal(i0["iter"], i1["iter"]) and np.array_equal(i0["status"], i1["status"])


def test_hand_back_when_no_constraint_is_active():
    b = problems.chain_mass(2, N=6, seed=2)
    for k in range(b.shape.N + 1):
        b.layout.view(b.qp, "dmask", k)[1] = 0.0
    o = default_opts()
    sol, info, redo = ob.fast_emul_solve(b, o, g=8)
    assert redo.tolist() == [1] and info["status"][1] == 100 and info["status"][0] == 0
