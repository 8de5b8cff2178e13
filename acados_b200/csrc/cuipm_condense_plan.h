// cuipm_condense_plan.h -- host-side construction of the condensing plan (block partition, condensed shape, index maps, layout
// tables) shared by the device wrapper (cuipm_condense.cu) and the host emulation used by the tests (oracle/condense_emul.cpp).
#ifndef CUIPM_CONDENSE_PLAN_H_
#define CUIPM_CONDENSE_PLAN_H_

#include <algorithm>
#include <vector>

#include "cuipm.h"
#include "cuipm_condense_core.h"

namespace cuipm_cond {

struct HostPlan
{
    int N = 0, N2 = 0;
    // condensed shape (arrays owned here)
    std::vector<int> nx2, nu2, nb2, ng2, ns2;
    std::vector<std::vector<int>> idxb2, rev2;
    std::vector<const int *> pidxb2, prev2;
    cuipm_shape cshape{};
    cuipm_layout *lo = nullptr, *lc = nullptr;
    // flat pools: every table of Plan is a slice of one of these (so that they can be uploaded in two copies)
    std::vector<int> ipool;
    std::vector<unsigned> upool;
    struct Slices { size_t o_dims[5], c_dims[5], o_ibp, o_ib, c_ibp, c_ib, blk_k0, blk_m, offu, offs, box_ptr, box_stage, box_i, gen_ptr, gen_stage, gen_kind, gen_i; size_t o_off[13], c_off[13]; size_t t_off; } sl{};
    unsigned t_stride = 0;                    // doubles per QP of the resident T_j buffer of the lhs / rhs split
    int nxmax = 0, n2max = 0;

    ~HostPlan() { cuipm_layout_destroy(lo); cuipm_layout_destroy(lc); }

    // Plan whose pointers refer to the given bases (host pools or their device copies)
    Plan plan(const int *ib, const unsigned *ub) const
    {
        Plan P{};
        auto lt = [&](const size_t dims[5], size_t ibp, size_t ibx, const size_t off[13], int Nn, const cuipm_layout *l) {
            LayoutTab t{};
            t.N = Nn;
            t.nx = ib + dims[0]; t.nu = ib + dims[1]; t.nb = ib + dims[2]; t.ng = ib + dims[3]; t.ns = ib + dims[4];
            t.idxb_ptr = ib + ibp; t.idxb = ib + ibx;
            const unsigned **f[13] = {&t.BAt, &t.RSQ, &t.DCt, &t.b, &t.rq, &t.d, &t.dmask, &t.Z, &t.z, &t.ux, &t.pi, &t.lam, &t.t};
            for (int i = 0; i < 13; i++) *f[i] = ub + off[i];
            t.qp_stride = (unsigned) l->qp_stride; t.sol_stride = (unsigned) l->sol_stride;
            return t;
        };
        P.o = lt(sl.o_dims, sl.o_ibp, sl.o_ib, sl.o_off, N, lo);
        P.c = lt(sl.c_dims, sl.c_ibp, sl.c_ib, sl.c_off, N2, lc);
        P.blk_k0 = ib + sl.blk_k0; P.blk_m = ib + sl.blk_m; P.st_offu = ib + sl.offu; P.st_offs = ib + sl.offs;
        P.box_ptr = ib + sl.box_ptr; P.box_stage = ib + sl.box_stage; P.box_i = ib + sl.box_i;
        P.gen_ptr = ib + sl.gen_ptr; P.gen_stage = ib + sl.gen_stage; P.gen_kind = ib + sl.gen_kind; P.gen_i = ib + sl.gen_i;
        P.nxmax = nxmax; P.n2max = n2max;
        P.t_off = ub + sl.t_off; P.t_stride = t_stride;
        return P;
    }
};

// block sizes as d_part_cond_qp_compute_block_size (external/hpipm/cond/x_part_cond.c:45-63)
inline bool build_plan(const cuipm_shape *sh, int cond_N, HostPlan &hp)
{
    const int N = sh->N;
    if (cond_N < 1 || cond_N > N) return false;
    hp.N = N; hp.N2 = cond_N;
    const int n1 = N / cond_N, r1 = N - cond_N * n1;
    std::vector<int> blk_k0, blk_m, offu(N + 1, 0), offs(N + 1, 0), box_ptr{0}, box_stage, box_i, gen_ptr{0}, gen_stage, gen_kind, gen_i;
    int k = 0;
    hp.idxb2.assign(cond_N + 1, {}); hp.rev2.assign(cond_N + 1, {});
    auto rev_of = [&](int j, int pos) { return (sh->ns[j] > 0 && sh->idxs_rev) ? sh->idxs_rev[j][pos] : -1; };
    for (int b = 0; b < cond_N; b++)
    {
        const int m = b < r1 ? n1 + 1 : n1, k0 = k;
        blk_k0.push_back(k0); blk_m.push_back(m);
        int ou = 0, os = 0;
        for (int j = k0; j < k0 + m; j++) { offu[j] = ou; offs[j] = os; ou += sh->nu[j]; os += sh->ns[j]; }
        const int nuu = ou;
        std::vector<int> grev;
        for (int j = k0; j < k0 + m; j++)
        {
            for (int i = 0; i < sh->nb[j]; i++)
            {
                const int var = sh->idxb[j][i], r = rev_of(j, i), r2 = r >= 0 ? offs[j] + r : -1;
                if (var < sh->nu[j]) { box_stage.push_back(j); box_i.push_back(i); hp.idxb2[b].push_back(offu[j] + var); hp.rev2[b].push_back(r2); }
                else if (j == k0) { box_stage.push_back(j); box_i.push_back(i); hp.idxb2[b].push_back(nuu + var - sh->nu[j]); hp.rev2[b].push_back(r2); }
                else { gen_stage.push_back(j); gen_kind.push_back(0); gen_i.push_back(i); grev.push_back(r2); }
            }
            for (int g = 0; g < sh->ng[j]; g++)
            {
                const int r = rev_of(j, sh->nb[j] + g);
                gen_stage.push_back(j); gen_kind.push_back(1); gen_i.push_back(g); grev.push_back(r >= 0 ? offs[j] + r : -1);
            }
        }
        box_ptr.push_back((int) box_stage.size()); gen_ptr.push_back((int) gen_stage.size());
        hp.nx2.push_back(sh->nx[k0]); hp.nu2.push_back(nuu); hp.nb2.push_back(box_ptr[b + 1] - box_ptr[b]);
        hp.ng2.push_back(gen_ptr[b + 1] - gen_ptr[b]); hp.ns2.push_back(os);
        hp.rev2[b].insert(hp.rev2[b].end(), grev.begin(), grev.end());
        hp.n2max = std::max(hp.n2max, nuu + sh->nx[k0]);
        k += m;
    }
    hp.nx2.push_back(sh->nx[N]); hp.nu2.push_back(sh->nu[N]); hp.nb2.push_back(sh->nb[N]); hp.ng2.push_back(sh->ng[N]); hp.ns2.push_back(sh->ns[N]);
    hp.idxb2[cond_N].assign(sh->idxb[N], sh->idxb[N] + sh->nb[N]);
    for (int i = 0; i < sh->nb[N] + sh->ng[N]; i++) hp.rev2[cond_N].push_back(rev_of(N, i));
    // the scratch blocks T, T2 (nxmax x n2max) also hold S_j T (nu_j x n2): size them by the larger of nx and nu
    for (int kk = 0; kk <= N; kk++) hp.nxmax = std::max(hp.nxmax, std::max(sh->nx[kk], sh->nu[kk]));
    for (int b = 0; b <= cond_N; b++)
    {
        if (hp.idxb2[b].empty()) hp.idxb2[b].push_back(0);
        if (hp.rev2[b].empty()) hp.rev2[b].push_back(-1);
    }
    for (int b = 0; b <= cond_N; b++) { hp.pidxb2.push_back(hp.idxb2[b].data()); hp.prev2.push_back(hp.rev2[b].data()); }
    hp.cshape = cuipm_shape{cond_N, hp.nx2.data(), hp.nu2.data(), hp.nb2.data(), hp.ng2.data(), hp.ns2.data(), hp.pidxb2.data(), hp.prev2.data()};
    hp.lo = cuipm_layout_create(sh);
    hp.lc = cuipm_layout_create(&hp.cshape);
    if (hp.lo->qp_stride >= ((size_t) 1 << 32) || hp.lc->qp_stride >= ((size_t) 1 << 32)) return false;
    // pools
    auto puti = [&](const int *p, size_t n) { size_t o = hp.ipool.size(); hp.ipool.insert(hp.ipool.end(), p, p + n); return o; };
    auto putv = [&](const std::vector<int> &v) { return puti(v.data(), v.size()); };
    auto dims = [&](const cuipm_shape *s, size_t out[5], size_t &ibp, size_t &ibx) {
        const int n = s->N + 1;
        out[0] = puti(s->nx, n); out[1] = puti(s->nu, n); out[2] = puti(s->nb, n); out[3] = puti(s->ng, n); out[4] = puti(s->ns, n);
        std::vector<int> ptr{0}, pool;
        for (int kk = 0; kk < n; kk++) { pool.insert(pool.end(), s->idxb[kk], s->idxb[kk] + s->nb[kk]); ptr.push_back((int) pool.size()); }
        pool.push_back(0);
        ibp = putv(ptr); ibx = putv(pool);
    };
    dims(sh, hp.sl.o_dims, hp.sl.o_ibp, hp.sl.o_ib);
    dims(&hp.cshape, hp.sl.c_dims, hp.sl.c_ibp, hp.sl.c_ib);
    hp.sl.blk_k0 = putv(blk_k0); hp.sl.blk_m = putv(blk_m); hp.sl.offu = putv(offu); hp.sl.offs = putv(offs);
    box_stage.push_back(0); box_i.push_back(0); gen_stage.push_back(0); gen_kind.push_back(0); gen_i.push_back(0);
    hp.sl.box_ptr = putv(box_ptr); hp.sl.box_stage = putv(box_stage); hp.sl.box_i = putv(box_i);
    hp.sl.gen_ptr = putv(gen_ptr); hp.sl.gen_stage = putv(gen_stage); hp.sl.gen_kind = putv(gen_kind); hp.sl.gen_i = putv(gen_i);
    auto offs13 = [&](const cuipm_layout *l, int n, size_t out[13]) {
        const size_t *a[13] = {l->off_BAt, l->off_RSQ, l->off_DCt, l->off_b, l->off_rq, l->off_d, l->off_dmask, l->off_Z, l->off_z, l->off_ux, l->off_pi, l->off_lam, l->off_t};
        for (int i = 0; i < 13; i++)
        {
            out[i] = hp.upool.size();
            for (int kk = 0; kk <= n; kk++) hp.upool.push_back((unsigned) a[i][kk]);
        }
    };
    offs13(hp.lo, N, hp.sl.o_off);
    offs13(hp.lc, cond_N, hp.sl.c_off);
    // T_j of every original stage (nx_j x n2 of its block), kept per QP between the lhs and the rhs pass
    hp.sl.t_off = hp.upool.size();
    {
        size_t off = 0;
        for (int b = 0; b < cond_N; b++)
            for (int j = blk_k0[b]; j < blk_k0[b] + blk_m[b]; j++)
            {
                hp.upool.push_back((unsigned) off);
                off += (size_t) sh->nx[j] * (size_t) (hp.nu2[b] + hp.nx2[b]);
            }
        hp.upool.push_back((unsigned) off);
        hp.t_stride = (unsigned) ((off + 1) & ~(size_t) 1);
    }
    return true;
}

}  // namespace cuipm_cond
#endif
