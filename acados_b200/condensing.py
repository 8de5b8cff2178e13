"""Batched partial (block) condensing of OCP-QP records on the host, vectorised over the batch in numpy.

Reference: ``ocp_qp_partial_condensing`` (acados/ocp_qp/ocp_qp_partial_condensing.c:523-689) -> HPIPM
``d_part_cond_qp_cond`` / ``d_part_cond_qp_expand_sol`` (external/hpipm/cond/x_part_cond.c:410-866, x_cond_aux.c): the N stages
of a QP are grouped into N2 blocks (sizes as ``d_part_cond_qp_compute_block_size``, x_part_cond.c:45-63: the first
N - N2*floor(N/N2) blocks one stage longer), the states inside a block are eliminated through the dynamics

    x_{j+1} = A_j x_j + B_j u_j + b_j   =>   x_j = Phi_j x + Gam_j u2 + c_j ,   u2 = [u_{k0}; u_{k0+1}; ...]

and every block becomes ONE stage in (x_{k0}, u2): dense Hessian / gradient, dynamics (Phi, Gam, c) of the block end, input
bounds stay boxes, the state bounds of the block's first stage stay boxes, state bounds and general constraints of the inner
stages become general constraints in (x, u2) with their bounds shifted by c_j, slacks are carried over with their index maps.
The terminal stage N stays as it is.  After the solve the inner states follow from the dynamics and the inner multipliers pi_j
from the stationarity conditions, backwards.

Records in, records out (layouts of acados_b200.problems / include/cuipm.h): the interface a device kernel for this row will
have.  This is the reference's *host* module restated (in the reference, too, condensing runs on the host in front of the QP
solver); the arithmetic is not HPIPM's (it builds the condensed Hessian through a Cholesky-like recursion), so condensed data
agree with the reference's to round-off, not bit for bit.
"""
from __future__ import annotations

from typing import List

import numpy as np

from .problems import Layout, Shape


def block_sizes(N: int, N2: int) -> List[int]:
    """Stages per block (x_part_cond.c:45-63); the terminal stage is not part of any block."""
    if not 1 <= N2 <= N:
        raise ValueError("need 1 <= cond_N <= N")
    n1, r1 = N // N2, N - N2 * (N // N2)
    return [n1 + 1] * r1 + [n1] * (N2 - r1)


def _sym(view: np.ndarray) -> np.ndarray:
    """Full symmetric matrices from a record view of RSQ (the record references the lower triangle only)."""
    up = np.triu(view)                         # view[b, c, r] = RSQ[r, c]: r >= c is the upper triangle of the view
    return up + np.swapaxes(np.triu(view, 1), 1, 2)


class BlockCondenser:
    def __init__(self, shape: Shape, cond_N: int):
        self.shape, self.N, self.N2 = shape, shape.N, cond_N
        self.lay = Layout(shape)
        sizes = block_sizes(shape.N, cond_N)
        self.blocks, k = [], 0
        for sz in sizes:
            self.blocks.append(list(range(k, k + sz)))
            k += sz
        assert k == shape.N
        nx2, nu2, nb2, ng2, ns2, idxb2, rev2 = [], [], [], [], [], [], []
        self.maps = []
        for blk in self.blocks:
            k0 = blk[0]
            off_u, off_s, o, os_ = {}, {}, 0, 0
            for j in blk:
                off_u[j], off_s[j] = o, os_
                o += shape.nu[j]
                os_ += shape.ns[j]
            nuu, nss = o, os_
            box, gen = [], []          # box: (stage, bound index, condensed variable); gen: (stage, 'bx' | 'g', index)
            for j in blk:
                for i, var in enumerate(shape.idxb[j]):
                    if var < shape.nu[j]:
                        box.append((j, i, off_u[j] + var))
                    elif j == k0:
                        box.append((j, i, nuu + var - shape.nu[j]))
                    else:
                        gen.append((j, "bx", i))
                for g in range(shape.ng[j]):
                    gen.append((j, "g", g))
            nx2.append(shape.nx[k0]); nu2.append(nuu); nb2.append(len(box)); ng2.append(len(gen)); ns2.append(nss)
            idxb2.append([v for (_, _, v) in box])

            def rev_of(j, pos):
                r = shape.idxs_rev[j][pos] if shape.ns[j] > 0 else -1
                return off_s[j] + r if r >= 0 else -1
            rev2.append([rev_of(j, i) for (j, i, _) in box] + [rev_of(j, i if kind == "bx" else shape.nb[j] + i) for (j, kind, i) in gen])
            self.maps.append(dict(off_u=off_u, off_s=off_s, box=box, gen=gen))
        Nl = shape.N                   # terminal stage unchanged
        nx2.append(shape.nx[Nl]); nu2.append(shape.nu[Nl]); nb2.append(shape.nb[Nl]); ng2.append(shape.ng[Nl]); ns2.append(shape.ns[Nl])
        idxb2.append(list(shape.idxb[Nl])); rev2.append(list(shape.idxs_rev[Nl]))
        self.cshape = Shape(cond_N, nx2, nu2, nb2, ng2, ns2, idxb2, rev2)
        self.clay = Layout(self.cshape)

    # ---- per-stage views of a batch of records -----------------------------------------------------------------
    def _stage(self, qp, k):
        sh, L = self.shape, self.lay
        nu, nx, nb, ng, ns = sh.nu[k], sh.nx[k], sh.nb[k], sh.ng[k], sh.ns[k]
        H = _sym(L.view(qp, "RSQ", k))
        rq, d, m = L.view(qp, "rq", k), L.view(qp, "d", k), L.view(qp, "dmask", k)
        out = dict(nu=nu, nx=nx, nb=nb, ng=ng, ns=ns, R=H[:, :nu, :nu], S=H[:, :nu, nu:], Q=H[:, nu:, nu:], r=rq[:, :nu], q=rq[:, nu:],
                   d=d, m=m, Z=L.view(qp, "Z", k), z=L.view(qp, "z", k))
        if k < sh.N:
            BA = L.view(qp, "BAt", k)            # (batch, nx_next, nu+nx)
            out.update(B=BA[:, :, :nu], A=BA[:, :, nu:], b=L.view(qp, "b", k))
        if ng > 0:
            DC = L.view(qp, "DCt", k)            # (batch, ng, nu+nx)
            out.update(D=DC[:, :, :nu], C=DC[:, :, nu:])
        return out

    def _transition(self, qp, blk):
        """Phi_j (nx_j x nx), Gam_j (nx_j x nu2), c_j for j over the block and its end."""
        sh, nbatch = self.shape, qp.shape[0]
        k0 = blk[0]
        nx, nu2 = sh.nx[k0], sum(sh.nu[j] for j in blk)
        Phi = [np.broadcast_to(np.eye(nx), (nbatch, nx, nx)).copy()]
        Gam = [np.zeros((nbatch, nx, nu2))]
        c = [np.zeros((nbatch, nx))]
        o = 0
        for j in blk:
            st = self._stage(qp, j)
            G = st["A"] @ Gam[-1]
            G[:, :, o:o + st["nu"]] += st["B"]
            Phi.append(st["A"] @ Phi[-1]); Gam.append(G)
            c.append(np.einsum("bij,bj->bi", st["A"], c[-1]) + st["b"])
            o += st["nu"]
        return Phi, Gam, c

    # ---- condense ------------------------------------------------------------------------------------------------
    def condense(self, qp: np.ndarray) -> np.ndarray:
        sh, cl = self.shape, self.clay
        nbatch = qp.shape[0]
        out = cl.new_qp(nbatch)
        for b, blk in enumerate(self.blocks):
            mp = self.maps[b]
            nx, nuu = self.cshape.nx[b], self.cshape.nu[b]
            nb2, ng2, ns2 = self.cshape.nb[b], self.cshape.ng[b], self.cshape.ns[b]
            Phi, Gam, c = self._transition(qp, blk)
            H = np.zeros((nbatch, nuu + nx, nuu + nx))
            g = np.zeros((nbatch, nuu + nx))
            for jj, j in enumerate(blk):
                st = self._stage(qp, j)
                T = np.concatenate([Gam[jj], Phi[jj]], axis=2)          # x_j = T [u2; x] + c_j
                ou, nu = mp["off_u"][j], st["nu"]
                QT = st["Q"] @ T
                H += np.swapaxes(T, 1, 2) @ QT
                qc = np.einsum("bij,bj->bi", st["Q"], c[jj]) + st["q"]
                g += np.einsum("bji,bj->bi", T, qc)
                if nu > 0:
                    ST = st["S"] @ T                                       # nu x (nuu+nx)
                    H[:, ou:ou + nu, :] += ST
                    H[:, :, ou:ou + nu] += np.swapaxes(ST, 1, 2)
                    H[:, ou:ou + nu, ou:ou + nu] += st["R"]
                    g[:, ou:ou + nu] += np.einsum("bij,bj->bi", st["S"], c[jj]) + st["r"]
            Hv = cl.view(out, "RSQ", b)
            Hv[:] = 0.5 * (H + np.swapaxes(H, 1, 2))
            cl.view(out, "rq", b)[:] = g
            BA = cl.view(out, "BAt", b)                                   # dynamics to the next block: x+ = Phi x + Gam u2 + c
            BA[:, :, :nuu], BA[:, :, nuu:] = Gam[-1], Phi[-1]
            cl.view(out, "b", b)[:] = c[-1]
            d2, m2 = cl.view(out, "d", b), cl.view(out, "dmask", b)
            for p, (j, i, _) in enumerate(mp["box"]):
                st = self._stage(qp, j)
                nbj, ngj = st["nb"], st["ng"]
                d2[:, p], d2[:, nb2 + ng2 + p] = st["d"][:, i], st["d"][:, nbj + ngj + i]
                m2[:, p], m2[:, nb2 + ng2 + p] = st["m"][:, i], st["m"][:, nbj + ngj + i]
            if ng2 > 0:
                DC = cl.view(out, "DCt", b)                               # (batch, ng2, nuu+nx): row = constraint
                for p, (j, kind, i) in enumerate(mp["gen"]):
                    st = self._stage(qp, j)
                    jj = j - blk[0]
                    nbj, ngj, nu = st["nb"], st["ng"], st["nu"]
                    T = np.concatenate([Gam[jj], Phi[jj]], axis=2)
                    if kind == "bx":
                        xi = sh.idxb[j][i] - nu
                        row, shift, pos = T[:, xi, :], c[jj][:, xi], i
                    else:
                        row = np.einsum("bj,bjk->bk", st["C"][:, i, :], T)
                        row[:, mp["off_u"][j]:mp["off_u"][j] + nu] += st["D"][:, i, :]
                        shift, pos = np.einsum("bj,bj->b", st["C"][:, i, :], c[jj]), nbj + i
                    DC[:, p, :] = row
                    d2[:, nb2 + p] = st["d"][:, pos] - shift
                    d2[:, 2 * nb2 + ng2 + p] = st["d"][:, nbj + ngj + pos] + shift      # upper bounds are stored negated
                    m2[:, nb2 + p], m2[:, 2 * nb2 + ng2 + p] = st["m"][:, pos], st["m"][:, nbj + ngj + pos]
            if ns2 > 0:
                Z2, z2 = cl.view(out, "Z", b), cl.view(out, "z", b)
                for j in blk:
                    st = self._stage(qp, j)
                    ns, os_, nbg = st["ns"], mp["off_s"][j], st["nb"] + st["ng"]
                    for half in (0, 1):
                        Z2[:, half * ns2 + os_:half * ns2 + os_ + ns] = st["Z"][:, half * ns:(half + 1) * ns]
                        z2[:, half * ns2 + os_:half * ns2 + os_ + ns] = st["z"][:, half * ns:(half + 1) * ns]
                        d2[:, 2 * (nb2 + ng2) + half * ns2 + os_:2 * (nb2 + ng2) + half * ns2 + os_ + ns] = st["d"][:, 2 * nbg + half * ns:2 * nbg + (half + 1) * ns]
                        m2[:, 2 * (nb2 + ng2) + half * ns2 + os_:2 * (nb2 + ng2) + half * ns2 + os_ + ns] = st["m"][:, 2 * nbg + half * ns:2 * nbg + (half + 1) * ns]
        # terminal stage: copied
        for f in ("RSQ", "DCt", "rq", "d", "dmask", "Z", "z"):
            o, sz = self.lay.off[f][self.N], self.lay.size[f][self.N]
            o2 = cl.off[f][self.N2]
            out[:, o2:o2 + sz] = qp[:, o:o + sz]
        return out

    # ---- expand ----------------------------------------------------------------------------------------------------
    def expand(self, qp: np.ndarray, sol2: np.ndarray) -> np.ndarray:
        """Solution records of the condensed QP -> solution records of the original (N-stage) shape."""
        sh, L, cl = self.shape, self.lay, self.clay
        nbatch = qp.shape[0]
        sol = L.new_sol(nbatch)
        for b, blk in enumerate(self.blocks):
            mp = self.maps[b]
            nuu, nx = self.cshape.nu[b], self.cshape.nx[b]
            nb2, ng2, ns2 = self.cshape.nb[b], self.cshape.ng[b], self.cshape.ns[b]
            ux2, lam2, t2 = cl.view(sol2, "ux", b), cl.view(sol2, "lam", b), cl.view(sol2, "t", b)
            x = ux2[:, nuu:nuu + nx]
            for j in blk:
                st = self._stage(qp, j)
                nu, ns, nbj, ngj, os_ = st["nu"], st["ns"], st["nb"], st["ng"], mp["off_s"][j]
                u = ux2[:, mp["off_u"][j]:mp["off_u"][j] + nu]
                uxj = L.view(sol, "ux", j)
                uxj[:, :nu], uxj[:, nu:nu + st["nx"]] = u, x
                uxj[:, nu + st["nx"]:nu + st["nx"] + ns] = ux2[:, nuu + nx + os_:nuu + nx + os_ + ns]
                uxj[:, nu + st["nx"] + ns:nu + st["nx"] + 2 * ns] = ux2[:, nuu + nx + ns2 + os_:nuu + nx + ns2 + os_ + ns]
                lamj, tj = L.view(sol, "lam", j), L.view(sol, "t", j)
                nbg = nbj + ngj
                for half in (0, 1):
                    lamj[:, 2 * nbg + half * ns:2 * nbg + (half + 1) * ns] = lam2[:, 2 * (nb2 + ng2) + half * ns2 + os_:2 * (nb2 + ng2) + half * ns2 + os_ + ns]
                    tj[:, 2 * nbg + half * ns:2 * nbg + (half + 1) * ns] = t2[:, 2 * (nb2 + ng2) + half * ns2 + os_:2 * (nb2 + ng2) + half * ns2 + os_ + ns]
                x = np.einsum("bij,bj->bi", st["A"], x) + np.einsum("bij,bj->bi", st["B"], u) + st["b"]
            for p, (j, i, _) in enumerate(mp["box"]):
                nbg = sh.nb[j] + sh.ng[j]
                for src, dst in ((lam2, L.view(sol, "lam", j)), (t2, L.view(sol, "t", j))):
                    dst[:, i], dst[:, nbg + i] = src[:, p], src[:, nb2 + ng2 + p]
            for p, (j, kind, i) in enumerate(mp["gen"]):
                nbg = sh.nb[j] + sh.ng[j]
                pos = i if kind == "bx" else sh.nb[j] + i
                for src, dst in ((lam2, L.view(sol, "lam", j)), (t2, L.view(sol, "t", j))):
                    dst[:, pos], dst[:, nbg + pos] = src[:, nb2 + p], src[:, 2 * nb2 + ng2 + p]
            # multipliers of the dynamics: the block end from the condensed QP, the inner ones backwards from stationarity in x_j
            pi = cl.view(sol2, "pi", b)
            L.view(sol, "pi", blk[-1])[:] = pi
            for j in reversed(blk[1:]):
                st = self._stage(qp, j)
                uxj, lamj = L.view(sol, "ux", j), L.view(sol, "lam", j)
                nu, nxj, nbj, ngj = st["nu"], st["nx"], st["nb"], st["ng"]
                xj, uj = uxj[:, nu:nu + nxj], uxj[:, :nu]
                g = np.einsum("bij,bj->bi", st["Q"], xj) + st["q"] + np.einsum("bji,bj->bi", st["A"], pi)
                if nu > 0:
                    g += np.einsum("bji,bj->bi", st["S"], uj)
                dl = lamj[:, nbj + ngj:2 * (nbj + ngj)] - lamj[:, :nbj + ngj]
                for i, var in enumerate(sh.idxb[j]):
                    if var >= nu:
                        g[:, var - nu] += dl[:, i]
                if ngj > 0:
                    g += np.einsum("bgi,bg->bi", st["C"], dl[:, nbj:])
                pi = g
                L.view(sol, "pi", j - 1)[:] = pi
        oN, szN = L.sol_stage[self.N], L.sol_stride - L.sol_stage[self.N]
        o2 = cl.sol_stage[self.N2]
        sol[:, oN:oN + szN] = sol2[:, o2:o2 + szN]
        return sol
