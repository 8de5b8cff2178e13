"""Shapes, record layout and synthetic OCP-QP batches for the cuipm solver.

Host-side helpers (numpy only).  The record layout mirrors ``include/cuipm.h`` (``cuipm_layout``); the
conventions are HPIPM's ``struct d_ocp_qp`` (reference: external/hpipm/include/hpipm_d_ocp_qp.h:54-71,
setters external/hpipm/ocp_qp/x_ocp_qp.c:1035-1267): per stage ``BAt = [B'; A']``, ``RSQ = [R S'; S Q]``
(lower triangle referenced), ``DCt = [D'; C']``, ``d = [lb, lg, -ub, -ug, lls, lus]``.

Problem families:
  * ``mass_spring``  -- the reference's own fixture (examples/c/no_interface_examples/mass_spring_model/
    mass_spring_qp.c:57-138,215-480), nx=8 nu=3 N=15, after x0 elimination (BASELINE config 1).
  * ``chain_mass``   -- chain of masses nx=21 nu=3 N=40, input box + 4 one-sided soft state bounds
    (shape of examples/acados_python/chain_mass/main.py:134-176) (BASELINE config 2 / headline).
  * ``random_qp``    -- random stable dynamics, any per-stage dims incl. general/soft/masked constraints.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field
from typing import List, Optional, Sequence

import numpy as np


def _ev(n: int) -> int:
    return (n + 1) & ~1


@dataclass
class Shape:
    """Batch-wide problem shape (mirror of ``cuipm_shape``)."""
    N: int
    nx: List[int]
    nu: List[int]
    nb: List[int]
    ng: List[int]
    ns: List[int]
    idxb: List[List[int]]
    idxs_rev: List[List[int]]
    _keep: list = field(default_factory=list, repr=False)

    def __post_init__(self):
        N = self.N
        for name in ("nx", "nu", "nb", "ng", "ns", "idxb", "idxs_rev"):
            assert len(getattr(self, name)) == N + 1, name
        for k in range(N + 1):
            assert len(self.idxb[k]) == self.nb[k]
            assert len(self.idxs_rev[k]) == self.nb[k] + self.ng[k]
            assert all(0 <= i < self.nu[k] + self.nx[k] for i in self.idxb[k])
            assert all(-1 <= i < self.ns[k] for i in self.idxs_rev[k])

    def nv(self, k):
        return self.nu[k] + self.nx[k]

    def nc(self, k):
        return 2 * (self.nb[k] + self.ng[k] + self.ns[k])

    def nx_next(self, k):
        return self.nx[k + 1] if k < self.N else 0

    def as_ctypes(self):
        """Returns a ctypes ``cuipm_shape`` (keeps the backing arrays alive on self)."""
        N = self.N

        def iarr(v):
            a = (C.c_int * len(v))(*v)
            self._keep.append(a)
            return a

        class CShape(C.Structure):
            _fields_ = [("N", C.c_int), ("nx", C.POINTER(C.c_int)), ("nu", C.POINTER(C.c_int)),
                        ("nb", C.POINTER(C.c_int)), ("ng", C.POINTER(C.c_int)), ("ns", C.POINTER(C.c_int)),
                        ("idxb", C.POINTER(C.POINTER(C.c_int))), ("idxs_rev", C.POINTER(C.POINTER(C.c_int)))]

        pidx = (C.POINTER(C.c_int) * (N + 1))()
        prev = (C.POINTER(C.c_int) * (N + 1))()
        for k in range(N + 1):
            pidx[k] = C.cast(iarr(self.idxb[k] or [0]), C.POINTER(C.c_int))
            prev[k] = C.cast(iarr(self.idxs_rev[k] or [-1]), C.POINTER(C.c_int))
        self._keep += [pidx, prev]
        s = CShape(N, C.cast(iarr(self.nx), C.POINTER(C.c_int)), C.cast(iarr(self.nu), C.POINTER(C.c_int)),
                   C.cast(iarr(self.nb), C.POINTER(C.c_int)), C.cast(iarr(self.ng), C.POINTER(C.c_int)),
                   C.cast(iarr(self.ns), C.POINTER(C.c_int)), pidx, prev)
        self._keep.append(s)
        return s


_QP_FIELDS = ("BAt", "RSQ", "DCt", "b", "rq", "d", "dmask", "Z", "z")
_SOL_FIELDS = ("ux", "pi", "lam", "t")


class Layout:
    """Offsets (in doubles) of every per-stage array inside one QP record / one solution record."""

    def __init__(self, shape: Shape):
        self.shape = shape
        N = shape.N
        self.off = {f: [0] * (N + 1) for f in _QP_FIELDS + _SOL_FIELDS}
        self.size = {f: [0] * (N + 1) for f in _QP_FIELDS + _SOL_FIELDS}
        o = s = 0
        self.qp_stage, self.sol_stage = [], []
        for k in range(N + 1):
            n, nx1, nc, ns2 = shape.nv(k), shape.nx_next(k), shape.nc(k), 2 * shape.ns[k]
            self.qp_stage.append(o)
            for f, sz in (("BAt", n * nx1), ("RSQ", n * n), ("DCt", n * shape.ng[k]), ("b", nx1), ("rq", n),
                          ("d", nc), ("dmask", nc), ("Z", ns2), ("z", ns2)):
                self.off[f][k], self.size[f][k] = o, sz
                o += _ev(sz)
            self.sol_stage.append(s)
            for f, sz in (("ux", n + ns2), ("pi", nx1), ("lam", nc), ("t", nc)):
                self.off[f][k], self.size[f][k] = s, sz
                s += _ev(sz)
        self.qp_stage.append(o)
        self.sol_stage.append(s)
        self.qp_stride, self.sol_stride = o, s

    # ---- views -------------------------------------------------------------------------------------
    def view(self, rec: np.ndarray, name: str, k: int) -> np.ndarray:
        """View of array ``name`` of stage k for all records: shape (nbatch, size) or, for matrices,
        (nbatch, ncols, nrows) (the record stores them column-major)."""
        o, sz = self.off[name][k], self.size[name][k]
        v = rec[:, o:o + sz]
        n = self.shape.nv(k)
        if name == "BAt":
            return v.reshape(rec.shape[0], self.shape.nx_next(k), n)
        if name == "RSQ":
            return v.reshape(rec.shape[0], n, n)
        if name == "DCt":
            return v.reshape(rec.shape[0], self.shape.ng[k], n)
        return v

    def new_qp(self, nbatch: int) -> np.ndarray:
        qp = np.zeros((nbatch, self.qp_stride))
        for k in range(self.shape.N + 1):
            self.view(qp, "dmask", k)[:] = 1.0
        return qp

    def new_sol(self, nbatch: int) -> np.ndarray:
        return np.zeros((nbatch, self.sol_stride))

    def gather(self, rec: np.ndarray, name: str) -> np.ndarray:
        """All stages of a vector field concatenated: (nbatch, total)."""
        return np.concatenate([self.view(rec, name, k) for k in range(self.shape.N + 1)], axis=1)

    def u_traj(self, sol: np.ndarray) -> np.ndarray:
        return np.concatenate([self.view(sol, "ux", k)[:, :self.shape.nu[k]] for k in range(self.shape.N + 1)], axis=1)


@dataclass
class Batch:
    shape: Shape
    layout: Layout
    qp: np.ndarray  # (nbatch, qp_stride) float64, C-contiguous
    name: str = ""

    @property
    def nbatch(self):
        return self.qp.shape[0]


# ----------------------------------------------------------------------------------------------------
# helpers
# ----------------------------------------------------------------------------------------------------

def _expm(M: np.ndarray) -> np.ndarray:
    """Batched matrix exponential by scaling and squaring with a degree-12 Taylor kernel (M: (..., n, n))."""
    nrm = np.max(np.sum(np.abs(M), axis=-1), axis=-1)
    s = np.maximum(0, np.ceil(np.log2(np.maximum(nrm, 1e-16))).astype(int) + 2)
    smax = int(np.max(s)) if np.ndim(s) else int(s)
    A = M / (2.0 ** smax)
    n = M.shape[-1]
    E = np.broadcast_to(np.eye(n), M.shape).copy()
    T = E.copy()
    for i in range(1, 13):
        T = T @ A / i
        E = E + T
    for _ in range(smax):
        E = E @ E
    return E


def _set_cost(lay: Layout, qp, k, R=None, S=None, Q=None, r=None, q=None):
    sh = lay.shape
    nu, nx = sh.nu[k], sh.nx[k]
    H = lay.view(qp, "RSQ", k)  # (nb, col, row): H[:, j, i] = element (i, j)
    if R is not None and nu:
        H[:, :nu, :nu] = np.swapaxes(R, -1, -2)
    if Q is not None and nx:
        H[:, nu:, nu:] = np.swapaxes(Q, -1, -2)
    if S is not None and nu and nx:   # S: (nx, nu) block at rows nu.., cols 0..nu  (= hpipm S' with S nu x nx)
        H[:, :nu, nu:] = np.swapaxes(S, -1, -2)
        H[:, nu:, :nu] = S
    g = lay.view(qp, "rq", k)
    if r is not None and nu:
        g[:, :nu] = r
    if q is not None and nx:
        g[:, nu:] = q


def _set_dyn(lay: Layout, qp, k, A=None, B=None, b=None):
    sh = lay.shape
    nu, nx = sh.nu[k], sh.nx[k]
    M = lay.view(qp, "BAt", k)  # (nb, nx1, n): M[:, j, i] = BAt(i, j) = [B';A'](i,j) = B(j,i) / A(j, i-nu)
    if B is not None and nu:
        M[:, :, :nu] = B
    if A is not None and nx:
        M[:, :, nu:] = A
    if b is not None:
        lay.view(qp, "b", k)[:] = b


def _set_box(lay: Layout, qp, k, lb, ub, lb_mask=None, ub_mask=None):
    sh = lay.shape
    nb, ng = sh.nb[k], sh.ng[k]
    d = lay.view(qp, "d", k)
    d[:, :nb] = lb
    d[:, nb + ng:2 * nb + ng] = -np.asarray(ub)
    m = lay.view(qp, "dmask", k)
    if lb_mask is not None:
        m[:, :nb] = lb_mask
    if ub_mask is not None:
        m[:, nb + ng:2 * nb + ng] = ub_mask


# ----------------------------------------------------------------------------------------------------
# config 1: the reference's mass-spring fixture
# ----------------------------------------------------------------------------------------------------

def mass_spring_system(Ts: float, nx: int, nu: int):
    """Discrete-time mass-spring chain (restates mass_spring_qp.c:57-138)."""
    pp = nx // 2
    T = -2 * np.eye(pp) + np.eye(pp, k=1) + np.eye(pp, k=-1)
    Ac = np.block([[np.zeros((pp, pp)), np.eye(pp)], [T, np.zeros((pp, pp))]])
    Bc = np.zeros((nx, nu))
    Bc[pp:pp + nu, :] = np.eye(nu)
    A = _expm(Ac * Ts)
    B = np.linalg.solve(Ac, (A - np.eye(nx)) @ Bc)
    return A, B


def mass_spring(nbatch: int = 1, nx: int = 8, nu: int = 3, N: int = 15, seed: Optional[int] = None,
                x0_scale: float = 0.0) -> Batch:
    """Config 1 (mass_spring_qp.c:215-480; dims of test/ocp_qp/test_qpsolvers.cpp:146-156) after the x0
    elimination the reference performs before calling the QP solver (d_ocp_qp_reduce_eq_dof,
    external/hpipm/ocp_qp/x_ocp_qp_red.c:278): stage 0 has nx=0 and b_0 = b + A x0.
    ``seed``: perturb x0 per instance by N(0, x0_scale) (instance 0 keeps the fixture's x0)."""
    nxs = [0] + [nx] * N
    nus = [nu] * N + [0]
    nbs = [nu] + [nu + nx] * (N - 1) + [nx]
    idxb = [list(range(nu))] + [list(range(nu + nx))] * (N - 1) + [list(range(nx))]
    shape = Shape(N, nxs, nus, nbs, [0] * (N + 1), [0] * (N + 1), idxb, [[-1] * n for n in nbs])
    lay = Layout(shape)
    qp = lay.new_qp(nbatch)
    A, B = mass_spring_system(0.5, nx, nu)
    b = np.full(nx, 0.1)
    x0 = np.zeros((nbatch, nx))
    x0[:, 0] = 2.5
    x0[:, 1] = 2.5
    if seed is not None and nbatch > 1:
        rng = np.random.default_rng(seed)
        x0[1:] += x0_scale * rng.standard_normal((nbatch - 1, nx))
    for k in range(N + 1):
        _set_cost(lay, qp, k, R=2 * np.eye(nu), S=np.zeros((nx, nu)), Q=np.eye(nx), r=0.2, q=0.1)
        if k < N:
            _set_dyn(lay, qp, k, A=A, B=B, b=(b + x0 @ A.T) if k == 0 else b)
        lb = np.concatenate([np.full(shape.nu[k], -0.5), np.full(shape.nx[k], -4.0)])
        _set_box(lay, qp, k, lb, -lb)
    return Batch(shape, lay, qp, f"mass_spring nx={nx} nu={nu} N={N}")


# ----------------------------------------------------------------------------------------------------
# config 2 / headline: chain of masses
# ----------------------------------------------------------------------------------------------------

def chain_mass(nbatch: int, n_mass: int = 5, N: int = 40, seed: int = 1234, Ts: float = 0.2,
               soft: bool = True, perturb: float = 0.1) -> Batch:
    """Chain of ``n_mass`` masses in 3-D: first mass fixed at the origin, ``n_mass-2`` free masses with
    position+velocity states, last mass position controlled through its velocity u (nx = 6(n_mass-2)+3,
    nu = 3; n_mass=5 gives nx=21).  Linear springs of zero rest length make the dynamics exactly linear;
    every instance has its own spring constants / masses (+-``perturb``), initial state and gradient, so all
    matrices differ across the batch (worst case of SURVEY 8(d)).  Constraints as in the reference's
    chain_mass example: |u|<=1 hard, the y-position of the free masses and of the end mass bounded from below
    by a wall, softened with slack penalties (one-sided: the upper bounds are masked out).  x0 is eliminated."""
    nf = n_mass - 2
    nx, nu = 6 * nf + 3, 3
    rng = np.random.default_rng(seed)
    # state ordering: [p_1..p_nf (3 each), p_end (3), v_1..v_nf (3 each)]
    npos = 3 * (nf + 1)
    D = 1.0 * (1 + perturb * rng.uniform(-1, 1, (nbatch, nf + 1)))     # spring k between mass k and k+1
    m = 0.033 * (1 + perturb * rng.uniform(-1, 1, (nbatch, nf))) * 30  # scaled masses
    Ac = np.zeros((nbatch, nx, nx))
    Bc = np.zeros((nbatch, nx, nu))
    for i in range(nf):
        Ac[:, 3 * i:3 * i + 3, npos + 3 * i:npos + 3 * i + 3] = np.eye(3)      # p_i' = v_i
        # m_i v_i' = D_i (p_{i-1} - p_i) + D_{i+1} (p_{i+1} - p_i), p_0 = 0
        r = slice(npos + 3 * i, npos + 3 * i + 3)
        Ac[:, r, 3 * i:3 * i + 3] -= ((D[:, i] + D[:, i + 1]) / m[:, i])[:, None, None] * np.eye(3)
        if i > 0:
            Ac[:, r, 3 * (i - 1):3 * i] += (D[:, i] / m[:, i])[:, None, None] * np.eye(3)
        Ac[:, r, 3 * (i + 1):3 * (i + 2)] += (D[:, i + 1] / m[:, i])[:, None, None] * np.eye(3)
        Ac[:, r, npos + 3 * i:npos + 3 * i + 3] -= 0.05 * np.eye(3)               # light damping
    Bc[:, 3 * nf:3 * nf + 3, :] = np.eye(3)                                     # p_end' = u
    # exact discretisation through the augmented exponential
    Maug = np.zeros((nbatch, nx + nu, nx + nu))
    Maug[:, :nx, :nx] = Ac * Ts
    Maug[:, :nx, nx:] = Bc * Ts
    E = _expm(Maug)
    A, B = E[:, :nx, :nx], E[:, :nx, nx:]
    g = np.zeros(nx)
    g[npos + 1::3] = -9.81 * 0.02                                                # gravity pull (y) on free masses
    b = np.broadcast_to(g * Ts, (nbatch, nx))

    nbx = nf + 1
    nsl = nbx if soft else 0
    ypos = [nu + 3 * i + 1 for i in range(nf + 1)]                              # y of free masses and end mass
    nxs = [0] + [nx] * N
    nus = [nu] * N + [0]
    nbs = [nu] + [nu + nbx] * (N - 1) + [nbx]
    nss = [0] + [nsl] * N
    idxb = [list(range(nu))] + [list(range(nu)) + ypos] * (N - 1) + [[i - nu for i in ypos] if nbx else []]
    srev = list(range(nbx)) if soft else [-1] * nbx
    rev = [[-1] * nu] + [[-1] * nu + srev] * (N - 1) + [srev]
    shape = Shape(N, nxs, nus, nbs, [0] * (N + 1), nss, idxb, rev)
    lay = Layout(shape)
    qp = lay.new_qp(nbatch)

    x0 = np.zeros((nbatch, nx))
    # stretched chain along x, perturbed, with some initial velocity
    for i in range(nf + 1):
        x0[:, 3 * i] = (i + 1) * 0.25
    x0[:, :npos] += 0.1 * rng.standard_normal((nbatch, npos))
    x0[:, npos:] += 0.5 * rng.standard_normal((nbatch, nx - npos))
    xref = np.zeros(nx)
    for i in range(nf + 1):
        xref[3 * i] = (i + 1) * 0.2
    Qd = np.concatenate([np.full(npos, 1.0), np.full(nx - npos, 0.25)])
    Qd[3 * nf:3 * nf + 3] = 25.0
    W = 0.05 * rng.standard_normal((nbatch, nx, 3))
    Q = np.eye(nx) * Qd + W @ np.swapaxes(W, 1, 2)
    R = np.eye(nu) * 0.01 * (1 + perturb * rng.uniform(-1, 1, (nbatch, 1, 1)))
    q = -(Q @ xref) + 0.02 * rng.standard_normal((nbatch, nx))
    r = 0.02 * rng.standard_normal((nbatch, nu))
    wall = (-0.1 if soft else -0.6) - 0.05 * rng.uniform(0, 1, (nbatch, 1))
    for k in range(N + 1):
        if k == 0:
            _set_cost(lay, qp, k, R=R, r=r)
            _set_dyn(lay, qp, k, B=B, b=b + np.einsum("bij,bj->bi", A, x0))
            _set_box(lay, qp, k, np.full(nu, -1.0), np.full(nu, 1.0))
            continue
        _set_cost(lay, qp, k, R=R, S=np.zeros((nx, nu)), Q=Q * (5.0 if k == N else 1.0), r=r,
                  q=q * (5.0 if k == N else 1.0))
        if k < N:
            _set_dyn(lay, qp, k, A=A, B=B, b=b)
        nuk = shape.nu[k]
        lb = np.concatenate([np.broadcast_to(np.full(nuk, -1.0), (nbatch, nuk)), np.broadcast_to(wall, (nbatch, nbx))], 1)
        ub = np.concatenate([np.full(nuk, 1.0), np.full(nbx, 1e9)])
        ubm = np.concatenate([np.ones(nuk), np.zeros(nbx)])
        _set_box(lay, qp, k, lb, ub, ub_mask=ubm)
        if nsl:
            lay.view(qp, "Z", k)[:] = 1e2
            lay.view(qp, "z", k)[:] = 1.0     # slack bounds lls = lus = 0 stay active
    return Batch(shape, lay, qp, f"chain_mass nx={nx} nu={nu} N={N}")


# ----------------------------------------------------------------------------------------------------
# random problems of arbitrary shape
# ----------------------------------------------------------------------------------------------------

def random_shape(N: int, nx: int, nu: int, nbu: Optional[int] = None, nbx: int = 0, ng: int = 0, ns: int = 0,
                 x0_eliminated: bool = True, terminal_nu: int = 0, seed: int = 0) -> Shape:
    """Uniform stage dims with the usual boundary stages (stage 0 nx=0 when x0 is eliminated; no input at N).
    ``ns`` soft constraints soften the last ns of the (state box + general) constraints of each stage."""
    rng = np.random.default_rng(seed)
    nbu = nu if nbu is None else nbu
    nxs = [0 if x0_eliminated else nx] + [nx] * N
    nus = [nu] * N + [terminal_nu]
    nbs, ngs, nss, idxb, rev = [], [], [], [], []
    for k in range(N + 1):
        bu = min(nbu, nus[k])
        bx = min(nbx, nxs[k])
        xi = sorted(rng.choice(nxs[k], bx, replace=False).tolist()) if bx else []
        ib = list(range(bu)) + [nus[k] + i for i in xi]
        g = ng if nxs[k] > 0 else 0
        s = min(ns, bx + g)
        r = [-1] * (len(ib) + g)
        for j in range(s):
            r[len(r) - 1 - j] = j
        nbs.append(len(ib)); ngs.append(g); nss.append(s); idxb.append(ib); rev.append(r)
    return Shape(N, nxs, nus, nbs, ngs, nss, idxb, rev)


def random_qp(shape: Shape, nbatch: int, seed: int = 0, umax: float = 1.0, xmax: float = 1.0, x0_scale: float = 1.0,
              mask_frac: float = 0.0, shared_matrices: bool = False) -> Batch:
    """Random strictly convex OCP-QPs (SURVEY 8(d) recipe): A = I + 0.1 G/sqrt(nx) rescaled to spectral radius
    <= 1.05, B ~ N/sqrt(nx), SPD cost with small S, boxes on inputs/states, general constraints on random
    combinations, slack penalties Z=1e2..1e3, z=1; optionally a fraction of upper bounds masked out."""
    rng = np.random.default_rng(seed)
    lay = Layout(shape)
    qp = lay.new_qp(nbatch)
    nb_ = 1 if shared_matrices else nbatch
    N = shape.N
    nxm = max(shape.nx)
    for k in range(N + 1):
        nu, nx, n = shape.nu[k], shape.nx[k], shape.nv(k)
        W = rng.standard_normal((nb_, n, n))
        H = np.eye(n) + 0.1 * W @ np.swapaxes(W, 1, 2) / max(n, 1)
        H[:, :nu, :nu] += np.eye(nu) * rng.uniform(0.01, 2.0, (nb_, 1, 1))
        lay.view(qp, "RSQ", k)[:] = H
        lay.view(qp, "rq", k)[:] = rng.uniform(-0.2, 0.2, (nbatch, n))
        if k < N:
            nx1 = shape.nx[k + 1]
            G = rng.standard_normal((nb_, nx1, nx))
            Ak = 0.1 * G / np.sqrt(max(nxm, 1))
            m = min(nx, nx1)
            Ak[:, np.arange(m), np.arange(m)] += 1.0
            if nx == nx1 and nx > 0:
                rho = np.max(np.abs(np.linalg.eigvals(Ak)), axis=-1)
                Ak *= np.minimum(1.0, 1.05 / rho)[:, None, None]
            Bk = rng.standard_normal((nb_, nx1, nu)) / np.sqrt(max(nxm, 1))
            bk = rng.uniform(-0.1, 0.1, (nbatch, nx1))
            if nx == 0:   # x0 eliminated: fold a random initial state into b_0
                A0 = np.eye(nx1) + 0.1 * rng.standard_normal((nb_, nx1, nx1)) / np.sqrt(nx1)
                bk = bk + np.einsum("bij,bj->bi", np.broadcast_to(A0, (nbatch, nx1, nx1)),
                                    x0_scale * rng.standard_normal((nbatch, nx1)))
            _set_dyn(lay, qp, k, A=Ak, B=Bk, b=bk)
        nb, ng, ns = shape.nb[k], shape.ng[k], shape.ns[k]
        if nb:
            isu = np.array(shape.idxb[k]) < nu
            hi = np.where(isu, umax, xmax) * rng.uniform(0.8, 1.2, (nbatch, nb))
            lo = -np.where(isu, umax, xmax) * rng.uniform(0.8, 1.2, (nbatch, nb))
            _set_box(lay, qp, k, lo, hi)
        if ng:
            Cg = rng.standard_normal((nb_, ng, n)) / np.sqrt(n)
            lay.view(qp, "DCt", k)[:] = Cg
            d = lay.view(qp, "d", k)
            d[:, nb:nb + ng] = -xmax * rng.uniform(0.8, 1.5, (nbatch, ng))
            d[:, 2 * nb + ng:2 * nb + 2 * ng] = -xmax * rng.uniform(0.8, 1.5, (nbatch, ng))
        if ns:
            lay.view(qp, "Z", k)[:] = rng.uniform(1e2, 1e3, (nbatch, 2 * ns))
            lay.view(qp, "z", k)[:] = 1.0
        if mask_frac > 0 and nb + ng:
            msk = lay.view(qp, "dmask", k)
            drop = rng.uniform(size=(1, nb + ng)) < mask_frac     # structural: same constraints for the whole batch
            msk[:, nb + ng:2 * (nb + ng)] = np.where(drop, 0.0, 1.0)
            rv = np.array(shape.idxs_rev[k])
            for i in np.nonzero(drop[0])[0]:
                if rv[i] >= 0:
                    msk[:, 2 * (nb + ng) + ns + rv[i]] = 0.0
    return Batch(shape, lay, qp, f"random N={N} nx={max(shape.nx)} nu={max(shape.nu)}")


def named_config(name: str, nbatch: Optional[int] = None, seed: int = 1234) -> Batch:
    """The BASELINE.json configs by name: c1 mass-spring, c2 chain-mass (headline shape), c3 pendulum-sized,
    c4 quadrotor-sized (uncondensed), c5 legged-sized."""
    if name == "c1":
        return mass_spring(nbatch or 1, seed=seed, x0_scale=0.5)
    if name == "c2":
        return chain_mass(nbatch or 4096, seed=seed)
    if name == "c3":
        sh = random_shape(20, 4, 1, nbx=0)
        return random_qp(sh, nbatch or 16384, seed=seed, umax=0.5, x0_scale=1.0)
    if name == "c4":
        sh = random_shape(50, 12, 4, nbx=6, ns=6)
        return random_qp(sh, nbatch or 8192, seed=seed, umax=0.5, xmax=1.0, x0_scale=1.0)
    if name == "c5":
        sh = random_shape(30, 48, 12, nbx=12, ns=12)
        return random_qp(sh, nbatch or 1024, seed=seed, umax=0.5, xmax=1.0, x0_scale=1.0)
    raise ValueError(name)
