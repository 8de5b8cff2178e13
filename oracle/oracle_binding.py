"""TEST INFRASTRUCTURE ONLY -- ctypes access to the CPU checkers under oracle/.

``oracle_solve``  : our plain-C restatement (oracle/liboracle_ipm.so, built by ``make -C oracle port``).
``ref_solve``     : the unmodified reference (HPIPM+BLASFEO behind acados' qp_solver vtable) compiled into
                    oracle/_ref/ by ``make -C oracle ref`` where /root/reference exists; the prebuilt .so files
                    travel to the GPU box.
Only tests/, __graft_entry__.smoke() and bench.py (cpu_baseline / --impl reference) may import this module.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from acados_b200.binding import INFO_DTYPE, STAT_M, CuipmOpts
from acados_b200.problems import Batch

_HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_LIB = os.path.join(_HERE, "liboracle_ipm.so")
REF_LIB = os.path.join(_HERE, "_ref", "libref_harness.so")
_libs = {}


def _load(path):
    if path not in _libs:
        if not os.path.exists(path):
            raise RuntimeError(f"{path} missing: run `make -C oracle` (needs /root/reference for the _ref part)")
        _libs[path] = C.CDLL(path)
    return _libs[path]


def have_ref() -> bool:
    return os.path.exists(REF_LIB)


def oracle_solve(batch: Batch, opts: CuipmOpts, sol0=None, want_stat=False, nthreads=0):
    lib = _load(ORACLE_LIB)
    nb = batch.nbatch
    sol = batch.layout.new_sol(nb) if sol0 is None else np.ascontiguousarray(sol0).copy()
    info = np.zeros(nb, dtype=INFO_DTYPE)
    stat = np.zeros((nb, opts.stat_max + 1, STAT_M)) if want_stat else None
    lib.oracle_solve.restype = C.c_int
    rc = lib.oracle_solve(C.byref(batch.shape.as_ctypes()), C.c_int(nb), C.c_void_p(batch.qp.ctypes.data),
                          C.c_void_p(sol.ctypes.data), C.c_void_p(info.ctypes.data),
                          C.c_void_p(stat.ctypes.data if want_stat else None), C.byref(opts), C.c_int(nthreads))
    if rc != 0:
        raise RuntimeError(f"oracle_solve: unsupported options ({rc})")
    return (sol, info, stat) if want_stat else (sol, info)


def oracle_residuals(batch: Batch, sol):
    lib = _load(ORACLE_LIB)
    info = np.zeros(batch.nbatch, dtype=INFO_DTYPE)
    sol = np.ascontiguousarray(sol)
    lib.oracle_residuals(C.byref(batch.shape.as_ctypes()), C.c_int(batch.nbatch), C.c_void_p(batch.qp.ctypes.data),
                         C.c_void_p(sol.ctypes.data), C.c_void_p(info.ctypes.data))
    return info


def oracle_layout(shape):
    from acados_b200.binding import c_layout_as_dict
    lib = _load(ORACLE_LIB)
    lib.oracle_layout_create.restype = C.c_void_p
    p = lib.oracle_layout_create(C.byref(shape.as_ctypes()))
    d = c_layout_as_dict(p, shape.N)
    lib.oracle_layout_destroy(C.c_void_p(p))
    return d


def ref_solve(batch: Batch, opts: CuipmOpts, sol0=None, want_stat=False, nthreads=0, nrep=1):
    """Returns (sol, info[, stat], timing) with timing = dict(solve_s=max over threads of time inside the
    reference's evaluate(), wall_s, threads)."""
    lib = _load(REF_LIB)
    nb = batch.nbatch
    sol = batch.layout.new_sol(nb) if sol0 is None else np.ascontiguousarray(sol0).copy()
    info = np.zeros(nb, dtype=INFO_DTYPE)
    stat = np.zeros((nb, opts.stat_max + 1, STAT_M)) if want_stat else None
    ts, tw = C.c_double(0), C.c_double(0)
    lib.ref_num_procs.restype = C.c_int
    lib.ref_solve(C.byref(batch.shape.as_ctypes()), C.c_int(nb), C.c_void_p(batch.qp.ctypes.data),
                  C.c_void_p(sol.ctypes.data), C.c_void_p(info.ctypes.data),
                  C.c_void_p(stat.ctypes.data if want_stat else None), C.byref(opts), C.c_int(nthreads), C.c_int(nrep),
                  C.byref(ts), C.byref(tw))
    timing = {"solve_s": ts.value, "wall_s": tw.value,
              "threads": nthreads if nthreads > 0 else int(lib.ref_num_procs())}
    return (sol, info, stat, timing) if want_stat else (sol, info, timing)


def _solve_sens(lib, fn, batch, opts, seed, adjoint, nthreads, sol0):
    nb = batch.nbatch
    sol = batch.layout.new_sol(nb) if sol0 is None else np.ascontiguousarray(sol0).copy()
    info = np.zeros(nb, dtype=INFO_DTYPE)
    seed = np.ascontiguousarray(seed, dtype=np.float64)
    assert seed.shape == sol.shape
    sens = np.zeros_like(sol)
    fn.restype = C.c_int
    rc = fn(C.byref(batch.shape.as_ctypes()), C.c_int(nb), C.c_void_p(batch.qp.ctypes.data), C.c_void_p(sol.ctypes.data),
            C.c_void_p(info.ctypes.data), C.byref(opts), C.c_int(nthreads), C.c_void_p(seed.ctypes.data),
            C.c_void_p(sens.ctypes.data), C.c_int(1 if adjoint else 0))
    if rc != 0:
        raise RuntimeError(f"solve_sens failed ({rc})")
    return sol, info, sens


def oracle_solve_sens(batch: Batch, opts: CuipmOpts, seed, adjoint=False, nthreads=0, sol0=None):
    """Solve, then one forward / adjoint solution-sensitivity evaluation per QP; seed and the returned sens are records
    in the solution layout with (seed_g, seed_b, seed_d, seed_m) in the (ux, pi, lam, t) slots."""
    lib = _load(ORACLE_LIB)
    return _solve_sens(lib, lib.oracle_solve_sens, batch, opts, seed, adjoint, nthreads, sol0)


def ref_solve_sens(batch: Batch, opts: CuipmOpts, seed, adjoint=False, nthreads=1, sol0=None):
    lib = _load(REF_LIB)
    return _solve_sens(lib, lib.ref_solve_sens, batch, opts, seed, adjoint, nthreads, sol0)


def ref_solve_xcond(batch: Batch, idxe0, cond_N: int, opts: CuipmOpts):
    """The reference's complete QP path (equality elimination, partial condensing to cond_N stages, HPIPM, expansion) on
    records of the FULL shape (x0 a stage-0 equality).  Returns (sol, info) with info['status'] the acados return value."""
    lib = _load(REF_LIB)
    nb = batch.nbatch
    sol = batch.layout.new_sol(nb)
    info = np.zeros(nb, dtype=INFO_DTYPE)
    idx = (C.c_int * max(1, len(idxe0)))(*[int(i) for i in idxe0])
    lib.ref_solve_xcond.restype = C.c_int
    lib.ref_solve_xcond(C.byref(batch.shape.as_ctypes()), C.c_int(len(idxe0)), idx, C.c_int(cond_N), C.c_int(nb),
                        C.c_void_p(batch.qp.ctypes.data), C.c_void_p(sol.ctypes.data), C.c_void_p(info.ctypes.data), C.byref(opts))
    return sol, info


EMUL_LIB = os.path.join(_HERE, "libcondense_emul.so")


def emul_condense_split(batch: Batch, new_qp, cond_N: int, nthreads: int = 128):
    """lhs pass on ``batch`` then rhs pass on the records ``new_qp`` (same matrices, other vectors; None: lhs only) of the
    product's condensing kernel body, run sequentially on the host (oracle/condense_emul.cpp).  Returns the condensed records."""
    lib = _load(EMUL_LIB)
    qs, ss = C.c_size_t(0), C.c_size_t(0)
    if not lib.emul_condensed_strides(C.byref(batch.shape.as_ctypes()), C.c_int(cond_N), C.byref(qs), C.byref(ss)):
        raise ValueError("bad cond_N")
    out = np.zeros((batch.nbatch, qs.value))
    nq = None if new_qp is None else np.ascontiguousarray(new_qp)
    rc = lib.emul_condense_split(C.byref(batch.shape.as_ctypes()), C.c_int(cond_N), C.c_int(batch.nbatch), C.c_void_p(batch.qp.ctypes.data),
                                 C.c_void_p(nq.ctypes.data if nq is not None else None), C.c_void_p(out.ctypes.data), C.c_int(nthreads))
    assert rc == 0
    return out


def emul_condense(batch: Batch, cond_N: int, nthreads: int = 128):
    """The product's block-condensing kernel body run sequentially on the host (oracle/condense_emul.cpp): returns the
    condensed records.  ``nthreads`` is the emulated CTA size (results must not depend on it)."""
    lib = _load(EMUL_LIB)
    qs, ss = C.c_size_t(0), C.c_size_t(0)
    if not lib.emul_condensed_strides(C.byref(batch.shape.as_ctypes()), C.c_int(cond_N), C.byref(qs), C.byref(ss)):
        raise ValueError("bad cond_N")
    out = np.zeros((batch.nbatch, qs.value))
    rc = lib.emul_condense(C.byref(batch.shape.as_ctypes()), C.c_int(cond_N), C.c_int(batch.nbatch), C.c_void_p(batch.qp.ctypes.data),
                           C.c_void_p(out.ctypes.data), C.c_int(nthreads))
    assert rc == 0
    return out


def emul_expand(batch: Batch, cond_N: int, sol2: np.ndarray, nthreads: int = 128):
    lib = _load(EMUL_LIB)
    sol2 = np.ascontiguousarray(sol2)
    out = batch.layout.new_sol(batch.nbatch)
    rc = lib.emul_expand(C.byref(batch.shape.as_ctypes()), C.c_int(cond_N), C.c_int(batch.nbatch), C.c_void_p(batch.qp.ctypes.data),
                         C.c_void_p(sol2.ctypes.data), C.c_void_p(out.ctypes.data), C.c_int(nthreads))
    assert rc == 0
    return out


FAST_EMUL_LIB = os.path.join(_HERE, "libfast_emul.so")


def fast_emul_solve(batch: Batch, opts: CuipmOpts, g: int = 8, order: int = 0, sol0=None, want_stat=False, rr: bool = False):
    """The product's throughput kernel body (acados_b200/csrc/cuipm_fast_core.h) executed on the host emulation of a
    warp (oracle/simt_emul.h).  Returns (sol, info[, stat], redo) with redo = indices of the QPs the kernel hands back
    to the generic kernel (cold paths)."""
    lib = _load(FAST_EMUL_LIB)
    nb = batch.nbatch
    sol = batch.layout.new_sol(nb) if sol0 is None else np.ascontiguousarray(sol0).copy()
    info = np.zeros(nb, dtype=INFO_DTYPE)
    stat = np.zeros((nb, opts.stat_max + 1, STAT_M)) if want_stat else None
    redo = np.zeros(nb + 1, dtype=np.int32)
    nredo = C.c_int(0)
    lib.fast_emul_solve.restype = C.c_int
    lib.fast_emul_set_rr(C.c_int(1 if rr else 0))      # iteration-sliced scheduling (rr_first / rr_loop) instead of one warp per QP group
    rc = lib.fast_emul_solve(C.byref(batch.shape.as_ctypes()), C.c_int(nb), C.c_void_p(batch.qp.ctypes.data),
                             C.c_void_p(sol.ctypes.data), C.c_void_p(info.ctypes.data),
                             C.c_void_p(stat.ctypes.data if want_stat else None), C.byref(opts), C.c_int(g), C.c_int(order),
                             C.c_void_p(redo.ctypes.data), C.byref(nredo))
    if rc != 0:
        raise RuntimeError(f"fast_emul_solve: rc={rc} (-1 shape not eligible, -2 no instance for this (nx, nu, g))")
    lib.fast_emul_set_rr(C.c_int(0))
    redo = np.sort(redo[:nredo.value])
    return (sol, info, stat, redo) if want_stat else (sol, info, redo)
