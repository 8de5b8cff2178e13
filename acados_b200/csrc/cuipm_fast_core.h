// cuipm_fast_core.h -- body of the throughput kernel of the batched OCP-QP interior-point solver.
//
// Mapping: a GROUP of G lanes owns one QP for the whole solve and a warp carries 32/G QPs in lock step, so that every
// serial piece of the algorithm (pivots, reciprocal square roots, the scalar logic of the IPM) is one instruction stream
// for 32/G QPs.  Lane l of a group owns rows l, l+G, l+2G, ... of the stage block; the rank-k updates are register tiles
// of (row slots) x 8 columns fed by one LDS per own row and 128-bit broadcast loads of the other operand.  Stage blocks
// are staged into shared memory with asynchronous copies (leading dimension LD = 2 mod 4: row and column accesses are
// both bank-conflict free, column starts 16-byte aligned).
//
// Algorithm and data layout are those of the generic kernel (cuipm_kernel.cu), which restates HPIPM's
// d_ocp_qp_ipm_solve (external/hpipm/ocp_qp/x_ocp_qp_ipm.c:2684-3120): same records, same work-record layout (the
// sensitivity kernel and the Riccati getters read what this kernel leaves behind).  Restrictions (checked on the host,
// cuipm_plan.h fast_plan): x0 eliminated, uniform interior stages, no general constraints.  Cold paths -- the LQ
// refactorisation, iterative refinement steps, a QP without active constraints -- are not here: a QP that needs one is
// handed back (status CUIPM_FAST_REDO, index appended to redo_list) and solved from scratch by the generic kernel.
//
// The file is written against a few warp primitives supplied by the including translation unit (FK_DEV, fk_lane,
// fk_sync, fk_shfl_xor, fk_any, fk_cp16, fk_cp8, fk_cp_wait, fk_ldg, fk_rsqrt, fk_atomic_inc): the CUDA instantiation is
// cuipm_fast.cu; oracle/fast_emul.cpp instantiates the same body on a host emulation of a warp for the CPU test-suite.
#ifndef CUIPM_FAST_CORE_H_
#define CUIPM_FAST_CORE_H_

#include "cuipm_device.h"

namespace cuipm {
namespace fastk {

FK_DEV int evn(int n) { return (n + 1) & ~1; }

// per-QP scalars of the IPM loop
struct QpState
{
    double mu, obj, gap, alpha, res_m_tau;
    double res_max[4];
};

template <int NX, int NU, int G>
struct Ker
{
    static constexpr int NM = NX + NU;                    // rows of an interior stage block
    static constexpr int QPW = 32 / G;                    // QPs per warp
    static constexpr int LD = ((NM + 2) / 4) * 4 + 2;     // >= NM+1, = 2 mod 4
    static constexpr int RPM = (NM + 1 + G - 1) / G;      // row slots per lane (rows incl. the gradient row)
    static constexpr int PADR = ((G * RPM > LD ? G * RPM - LD : 0) + 3) & ~1;
    static constexpr int SZA = LD * NX + PADR;            // [A; b'] / A Lxx / staged BAt
    static constexpr int SZL = LD * NM + PADR;            // L_{k+1} -> L_k (factorisation); L_{k+1} (substitutions)
    static constexpr int SZU = LD * NU + 2;               // first nu columns of L_k (substitutions)
    static constexpr int SZD = 16;                        // 4 x 4 diagonal block
    static constexpr int MATS = SZA + SZL + SZU + SZD;

    // stage kinds: 0 = first (nx = 0), 1 = interior, 2 = last (nu = 0)
    template <int KIND>
    struct KD
    {
        static constexpr int nx = KIND == 0 ? 0 : NX, nu = KIND == 2 ? 0 : NU, nx1 = KIND == 2 ? 0 : NX;
    };

    const FastArgs &A;
    int li, gq;                // lane within the group, group within the warp
    double *MA, *ML, *LU, *DD, *V;
    const double *qp;
    double *sol, *wk;
    bool act;                  // this group's QP is being solved: global stores enabled

    struct View { const double *q; double *s; double *w; const int *ip; };

    FK_DEV Ker(const FastArgs &a, double *smem) : A(a)
    {
        const int lane = fk_lane();
        li = lane % G;
        gq = lane / G;
        double *S = smem + (size_t) gq * a.gstride;
        MA = S; ML = MA + SZA; LU = ML + SZL; DD = LU + SZU; V = DD + SZD;
        qp = nullptr; sol = nullptr; wk = nullptr; act = false;
    }

    template <int KIND>
    FK_DEV const StageDesc &sdk() const { return KIND == 0 ? A.s0 : (KIND == 1 ? A.s1 : A.sN); }
    template <int KIND>
    FK_DEV View view(int k) const
    {
        const unsigned kk = KIND == 1 ? (unsigned) (k - 1) : 0u;
        return View{qp + kk * A.qs, sol + kk * A.ss, wk + kk * A.ws, A.ipool + (int) kk * A.is};
    }
    FK_DEV const StageDesc &sdr(int k) const { return k == 0 ? A.s0 : (k == A.N ? A.sN : A.s1); }
    FK_DEV View viewr(int k) const
    {
        const unsigned kk = (k >= 1 && k < A.N) ? (unsigned) (k - 1) : 0u;
        return View{qp + kk * A.qs, sol + kk * A.ss, wk + kk * A.ws, A.ipool + (int) kk * A.is};
    }

    // ---- group reductions ---------------------------------------------------------------------------
    FK_DEV double gsum(double v) const
    {
#pragma unroll
        for (int m = G / 2; m > 0; m >>= 1) v += fk_shfl_xor(v, m);
        return v;
    }
    FK_DEV double gmin(double v) const
    {
#pragma unroll
        for (int m = G / 2; m > 0; m >>= 1) v = fmin(v, fk_shfl_xor(v, m));
        return v;
    }
    // max of non-negative values; NaN is propagated (BLASFEO VECNRM_INF semantics, d_aux_lib4.c:4893-4995)
    FK_DEV double gmax_nan(double v, int isnan_) const
    {
#pragma unroll
        for (int m = G / 2; m > 0; m >>= 1)
        {
            v = fmax(v, fk_shfl_xor(v, m));
            isnan_ |= fk_shfl_xor_i(isnan_, m);
        }
        return isnan_ ? NAN : v;
    }
    FK_DEV void st(double *p, double v) const { if (act) *p = v; }

    // ---- staging ------------------------------------------------------------------------------------
    // column-major R x C block (ld R) of a record -> shared memory (ld LD), asynchronously; fk_cp_wait() + fk_sync() complete it
    template <int R, int C>
    FK_DEV void stage_mat(double *dst, const double *src) const
    {
        if (R % 2 == 0)
        {
            constexpr int H = R / 2 > 0 ? R / 2 : 1, T = H * C;
#pragma unroll 4
            for (int e = li; e < T; e += G)
            {
                const int c = e / H, h = e - c * H;
                fk_cp16(dst + LD * c + 2 * h, src + R * c + 2 * h);
            }
        }
        else
        {
            constexpr int T = R * C, R1 = R > 0 ? R : 1;
#pragma unroll 4
            for (int e = li; e < T; e += G)
            {
                const int c = e / R1, r = e - c * R1;
                fk_cp8(dst + LD * c + r, src + e);
            }
        }
    }
    // row i of the symmetric n x n matrix H of which the lower triangle is stored (column-major, ld n), times x (shared)
    template <int n>
    FK_DEV double gdot_sym(const double *H, int i, const double *x) const
    {
        double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
        int j = 0;
#pragma unroll 2
        for (; j + 3 < n; j += 4)
        {
            const double a0 = fk_ldg(H + (j <= i ? i + n * j : j + n * i));
            const double a1 = fk_ldg(H + (j + 1 <= i ? i + n * (j + 1) : j + 1 + n * i));
            const double a2 = fk_ldg(H + (j + 2 <= i ? i + n * (j + 2) : j + 2 + n * i));
            const double a3 = fk_ldg(H + (j + 3 <= i ? i + n * (j + 3) : j + 3 + n * i));
            s0 += a0 * x[j]; s1 += a1 * x[j + 1]; s2 += a2 * x[j + 2]; s3 += a3 * x[j + 3];
        }
        for (; j < n; j++) s0 += fk_ldg(H + (j <= i ? i + n * j : j + n * i)) * x[j];
        return (s0 + s1) + (s2 + s3);
    }

    // ---------------------------------------------------------------------------------------------
    // residuals of the QP at the iterate (OCP_QP_RES_COMPUTE, x_ocp_qp_res.c:345-531) -> residual set 0, with
    // UPDATE_VAR_QP fused (x_core_qp_ipm_aux.c:472-582: the iterate first moves by alpha_u along the step, with the
    // step shortening and the t/lam clipping) and the affine complementarity right-hand side of the next
    // iteration (res_m = lam*t - tau_min, backup lam*t; BACKUP_RES_M / COMPUTE_TAU_MIN_QP :672-781).
    // ---------------------------------------------------------------------------------------------
    struct ResAcc
    {
        double a_mu, a_obj, a_gap, m0, m1, m2, m3, m4;
        int f0, f1, f2, f3, f4;
    };

    template <int KIND>
    FK_DEV void res_stage(int k, int update, double alpha_u, ResAcc &R)
    {
        constexpr int nx = KD<KIND>::nx, nu = KD<KIND>::nu, n = nx + nu, nx1 = KD<KIND>::nx1;
        const StageDesc &sd = sdk<KIND>();
        const View v = view<KIND>(k);
        const int nb = sd.nb, ns = sd.ns, nc = sd.nc;
        const int *idxb = v.ip + sd.idx_off, *rev = idxb + nb;
        double *ux = V, *x1 = ux + A.nve, *pi = x1 + evn(NX), *pim = pi + evn(NX);
        double *lam = pim + evn(NX), *lamr = lam + A.nce, *t = lamr + A.nce, *msk = t + A.nce;
        double *tmp0 = msk + A.nce, *tmp1 = tmp0 + A.nbe, *g_ = tmp1 + A.nbe;
        if (nx1 > 0) stage_mat<n, nx1>(MA, v.q + sd.q_BAt);
        // ---- vectors of this stage (optionally moved along the step) to shared memory
        {
            double *gu = v.s + sd.sol.ux;
            const double *du = v.w + sd.step.ux;
            for (int i = li; i < n + 2 * ns; i += G)
            {
                double x = gu[i];
                if (update) { x += alpha_u * du[i]; st(gu + i, x); }
                ux[i] = x;
            }
        }
        if (nx1 > 0)
        {
            const StageDesc &s1 = sdr(k + 1);
            const View v1 = viewr(k + 1);
            const double *gu1 = v1.s + s1.sol.ux + s1.nu, *du1 = v1.w + s1.step.ux + s1.nu, *dp = v.w + sd.step.pi;
            double *gp = v.s + sd.sol.pi;
            for (int j = li; j < nx1; j += G)
            {
                double x = gu1[j], p = gp[j];
                if (update) { x += alpha_u * du1[j]; p += alpha_u * dp[j]; st(gp + j, p); }
                x1[j] = x;
                pi[j] = p;
            }
        }
        {
            double *gl = v.s + sd.sol.lam, *gt = v.s + sd.sol.t;
            const double *gm = v.q + sd.q_dmask, *dl = v.w + sd.step.lam, *dtt = v.w + sd.step.t;
            double *bl = wk + A.w_bkp + (sd.sol.lam + (KIND == 1 ? (unsigned) (k - 1) * A.ss : 0u));
            double *bt = wk + A.w_bkp + (sd.sol.t + (KIND == 1 ? (unsigned) (k - 1) * A.ss : 0u));
            for (int i = li; i < nc; i += G)
            {
                double l = gl[i], tt = gt[i];
                const double mk = fk_ldg(gm + i);
                if (update)
                {
                    // iterate of the factorisation just used (UPDATE_VAR_QP backups, x_core_qp_ipm_aux.c:534-575): the point the
                    // sensitivities are evaluated at
                    st(bl + i, l);
                    st(bt + i, tt);
                    l += alpha_u * dl[i];
                    tt += alpha_u * dtt[i];
                    if (A.o.t_lam_min == 2)
                    {
                        l = l <= A.o.lam_min ? A.o.lam_min : l;
                        tt = tt <= A.o.t_min ? A.o.t_min : tt;
                    }
                    l *= mk;
                    st(gl + i, l);
                    st(gt + i, tt);
                }
                lamr[i] = l;
                lam[i] = l * mk;
                t[i] = tt;
                msk[i] = mk;
            }
        }
        fk_cp_wait();
        fk_sync();
        for (int i = li; i < nb; i += G) tmp0[i] = lam[nb + i] - lam[i];
        fk_sync();
        // ---- rows of res_g (lane = row), res_b (lane = column)
        const double *Hg = v.q + sd.q_RSQ;
        {
            const double *gvec = v.q + sd.q_rq;
            for (int i = li; i < n; i += G)
            {
                const double acc = gdot_sym<n>(Hg, i, ux);
                const double gv = fk_ldg(gvec + i);
                double r = acc + 2.0 * gv;
                R.a_obj += 0.5 * r * ux[i];
                r -= gv;
                R.a_gap += r * ux[i];
                if (nx > 0 && i >= nu) r -= pim[i - nu];
                if (nx1 > 0)
                {
                    double s0 = 0.0, s1 = 0.0;
                    const double *arow = MA + i;
                    int j = 0;
#pragma unroll 4
                    for (; j + 1 < nx1; j += 2) { s0 += arow[LD * j] * pi[j]; s1 += arow[LD * (j + 1)] * pi[j + 1]; }
                    if (j < nx1) s0 += arow[LD * j] * pi[j];
                    r += s0 + s1;
                }
                g_[i] = r;
            }
            if (nx1 > 0)
            {
                const double *bvec = v.q + sd.q_b;
                double *ob = v.w + sd.res.b;
                for (int j = li; j < nx1; j += G)
                {
                    double s0 = 0.0, s1 = 0.0;
                    const double *acol = MA + LD * j;
                    int i = 0;
#pragma unroll 4
                    for (; i + 1 < n; i += 2) { s0 += acol[i] * ux[i]; s1 += acol[i + 1] * ux[i + 1]; }
                    if (i < n) s0 += acol[i] * ux[i];
                    const double bv = fk_ldg(bvec + j);
                    const double r = bv - x1[j] + (s0 + s1);
                    st(ob + j, r);
                    const double a = fabs(r);
                    R.m1 = fmax(R.m1, a);
                    R.f1 |= (a != a);
                    R.a_gap -= bv * pi[j];
                }
            }
        }
        fk_sync();
        // ---- box scatter, slack rows
        for (int i = li; i < nb; i += G)
        {
            const int ix = idxb[i];
            tmp1[i] = ux[ix];
            g_[ix] += tmp0[i];
        }
        if (ns > 0)
        {
            const double *Z = v.q + sd.q_Z, *zvec = v.q + sd.q_z;
            for (int j = li; j < 2 * ns; j += G)
            {
                const double sj = ux[n + j], zz = fk_ldg(zvec + j);
                double r = fk_ldg(Z + j) * sj + 2.0 * zz;
                R.a_obj += 0.5 * r * sj;
                r -= zz;
                R.a_gap += r * sj;
                r -= lam[2 * nb + j];
                const int jj = j < ns ? j : j - ns, offl = j < ns ? 0 : nb;
                for (int i = 0; i < nb; i++)
                    if (rev[i] == jj) r -= lam[offl + i];
                g_[n + j] = r;
            }
        }
        fk_sync();
        // ---- res_d, res_m
        {
            const double *dvec = v.q + sd.q_d;
            double *od = v.w + sd.res.d, *om = v.w + sd.res.m, *obk = v.w + sd.w_rmb;
            for (int i = li; i < nc; i += G)
            {
                const double dv = fk_ldg(dvec + i);
                double r;
                if (i < 2 * nb)
                {
                    const int up = i >= nb, ii = up ? i - nb : i;
                    const double x = tmp1[ii];
                    r = t[i] + dv + (up ? x : -x);
                    if (ns > 0 && rev[ii] >= 0) r -= ux[n + (up ? ns : 0) + rev[ii]];
                }
                else
                    r = t[i] - ux[n + (i - 2 * nb)] + dv;
                r *= msk[i];
                st(od + i, r);
                double a = fabs(r);
                R.m2 = fmax(R.m2, a);
                R.f2 |= (a != a);
                R.a_gap -= dv * lam[i];
                double mm = lam[i] * t[i];
                mm *= msk[i];
                R.a_mu += fabs(mm);
                st(obk + i, mm);
                double ma = mm - A.o.tau_min;
                ma *= msk[i];
                st(om + i, ma);                                  // affine rhs of the next iteration
                const double a4 = fabs(mm - A.o.tau_min * msk[i]);
                R.m4 = fmax(R.m4, a4);
                R.f4 |= (a4 != a4);
                a = fabs(mm);
                R.m3 = fmax(R.m3, a);
                R.f3 |= (a != a);
            }
            double *og = v.w + sd.res.g;
            for (int i = li; i < n + 2 * ns; i += G)
            {
                const double r = g_[i];
                st(og + i, r);
                const double a = fabs(r);
                R.m0 = fmax(R.m0, a);
                R.f0 |= (a != a);
            }
        }
        fk_sync();
        for (int j = li; j < nx1; j += G) pim[j] = pi[j];      // pi_k is "pi_{k-1}" of the next stage
        fk_sync();
    }

    FK_DEV void res_pass(int update, double alpha_u, QpState &Q)
    {
        ResAcc R;
        R.a_mu = R.a_obj = R.a_gap = R.m0 = R.m1 = R.m2 = R.m3 = R.m4 = 0.0;
        R.f0 = R.f1 = R.f2 = R.f3 = R.f4 = 0;
        if (update && alpha_u < 1.0) alpha_u = alpha_u * ((1.0 - alpha_u) * 0.99 + alpha_u * 0.9999999);
        fk_sync();
        res_stage<0>(0, update, alpha_u, R);
        for (int k = 1; k < A.N; k++) res_stage<1>(k, update, alpha_u, R);
        res_stage<2>(A.N, update, alpha_u, R);
        Q.res_max[0] = gmax_nan(R.m0, R.f0);
        Q.res_max[1] = gmax_nan(R.m1, R.f1);
        Q.res_max[2] = gmax_nan(R.m2, R.f2);
        Q.res_max[3] = gmax_nan(R.m3, R.f3);
        Q.mu = gsum(R.a_mu) * nc_mask_inv;
        Q.obj = gsum(R.a_obj);
        Q.gap = gsum(R.a_gap);
        Q.res_m_tau = gmax_nan(R.m4, R.f4);
    }

    // ---------------------------------------------------------------------------------------------
    // slack elimination (x_ocp_qp_kkt.c:220-335, 431-520): tmp0/tmp1 = effective Gamma / gamma of the
    // softened constraints; ds = slack part of the step rhs; Zi = inverse of the slack Hessian.
    // ---------------------------------------------------------------------------------------------
    FK_DEV void cond_slacks(int nb, int ns, const int *rev, const double *Z, int fact, const double *Gam, const double *gam,
                            const double *rgs, double *Zi, double *ds, double *tmp0, double *tmp1) const
    {
        for (int j = li; j < 2 * ns; j += G)
        {
            const int jj = j < ns ? j : j - ns, offc = j < ns ? 0 : nb;
            double zi = 0.0, d = rgs[j] + gam[2 * nb + j];
            if (fact) zi = fk_ldg(Z + j) + A.o.reg_prim + Gam[2 * nb + j];
            for (int i = 0; i < nb; i++)
                if (rev[i] == jj)
                {
                    if (fact) zi += Gam[offc + i];
                    d += gam[offc + i];
                }
            if (fact) Zi[j] = 1.0 / zi;
            ds[j] = d;
        }
        fk_sync();
        for (int i = li; i < nb; i += G)
        {
            const int j = rev[i];
            double t0l, t0u, t1l, t1u;
            if (j != -1)
            {
                t0l = Gam[i] - Gam[i] * Zi[j] * Gam[i];
                t0u = Gam[nb + i] - Gam[nb + i] * Zi[ns + j] * Gam[nb + i];
                t1l = gam[i] - Gam[i] * Zi[j] * ds[j];
                t1u = gam[nb + i] - Gam[nb + i] * Zi[ns + j] * ds[ns + j];
            }
            else
            {
                t0l = Gam[i]; t0u = Gam[nb + i]; t1l = gam[i]; t1u = gam[nb + i];
            }
            if (fact) tmp0[i] = t0l + t0u;
            tmp1[i] = t1l - t1u;
        }
    }

    // ---------------------------------------------------------------------------------------------
    // 4-column panel of the left-looking Cholesky: x[m][0..3] hold the raw (updated) entries of columns j0..j0+3 of
    // this lane's rows; the 4 x 4 diagonal block is published through DD, factorised redundantly by every lane (pivot
    // rule blasfeo_ref/x_lapack_ref.c:84-91: a non-positive pivot gives a zero column), the rows below are scaled.
    // Results stay in x, go to ML (final columns of L) and to the work record (lower part; row n -> lrow).
    // ---------------------------------------------------------------------------------------------
    template <int n, int RP, int W4>
    FK_DEV void panel4(int j0, int m0, double (&x)[RPM][8], int xo, double *Lg, double *lrow, double *Linv)
    {
#pragma unroll
        for (int m = 0; m < RP; m++)
        {
            if (m < m0) continue;
            const int rr = li + G * m - j0;
            if (rr >= 0 && rr < W4)
            {
#pragma unroll
                for (int q = 0; q < W4; q++) DD[rr + 4 * q] = x[m][xo + q];
            }
        }
        fk_sync();
        double d00 = DD[0], d10 = 0, d20 = 0, d30 = 0, d11 = 0, d21 = 0, d31 = 0, d22 = 0, d32 = 0, d33 = 0;
        if (W4 > 1) { d10 = DD[1]; d11 = DD[5]; }
        if (W4 > 2) { d20 = DD[2]; d21 = DD[6]; d22 = DD[10]; }
        if (W4 > 3) { d30 = DD[3]; d31 = DD[7]; d32 = DD[11]; d33 = DD[15]; }
        const double i0 = d00 > 0.0 ? fk_rsqrt(d00) : 0.0;
        const double l10 = d10 * i0, l20 = d20 * i0, l30 = d30 * i0;
        d11 -= l10 * l10;
        const double i1 = d11 > 0.0 ? fk_rsqrt(d11) : 0.0;
        const double l21 = (d21 - l20 * l10) * i1, l31 = (d31 - l30 * l10) * i1;
        d22 -= l20 * l20 + l21 * l21;
        const double i2 = d22 > 0.0 ? fk_rsqrt(d22) : 0.0;
        const double l32 = (d32 - l30 * l20 - l31 * l21) * i2;
        d33 -= l30 * l30 + l31 * l31 + l32 * l32;
        const double i3 = d33 > 0.0 ? fk_rsqrt(d33) : 0.0;
#pragma unroll
        for (int m = 0; m < RP; m++)
        {
            if (m < m0) continue;
            const int r = li + G * m, rr = r - j0;     // position inside the panel: rows 0..3 form the diagonal block
            double x0 = x[m][xo] * i0, x1 = 0.0, x2 = 0.0, x3 = 0.0;
            if (rr == 0) x0 = d00 * i0;
            if (W4 > 1) x1 = rr == 0 ? 0.0 : (rr == 1 ? d11 * i1 : (x[m][xo + 1] - x0 * l10) * i1);
            if (W4 > 2) x2 = rr <= 1 ? 0.0 : (rr == 2 ? d22 * i2 : (x[m][xo + 2] - x0 * l20 - x1 * l21) * i2);
            if (W4 > 3) x3 = rr <= 2 ? 0.0 : (rr == 3 ? d33 * i3 : (x[m][xo + 3] - x0 * l30 - x1 * l31 - x2 * l32) * i3);
            x[m][xo] = x0;
            if (W4 > 1) x[m][xo + 1] = x1;
            if (W4 > 2) x[m][xo + 2] = x2;
            if (W4 > 3) x[m][xo + 3] = x3;
            if (rr >= 0 && r <= n)
            {
                double *mr = ML + r + LD * j0;
                mr[0] = x0;
                if (W4 > 1) mr[LD] = x1;
                if (W4 > 2) mr[2 * LD] = x2;
                if (W4 > 3) mr[3 * LD] = x3;
                if (act)
                {
                    if (r < n)
                    {
                        double *gr = Lg + r + n * j0;
                        gr[0] = x0;
                        if (W4 > 1 && rr >= 1) gr[n] = x1;
                        if (W4 > 2 && rr >= 2) gr[2 * n] = x2;
                        if (W4 > 3 && rr >= 3) gr[3 * n] = x3;
                    }
                    else
                    {
                        lrow[j0] = x0;
                        if (W4 > 1) lrow[j0 + 1] = x1;
                        if (W4 > 2) lrow[j0 + 2] = x2;
                        if (W4 > 3) lrow[j0 + 3] = x3;
                    }
                }
            }
        }
        if (li == 0)
        {
            Linv[j0] = i0;
            if (W4 > 1) Linv[j0 + 1] = i1;
            if (W4 > 2) Linv[j0 + 2] = i2;
            if (W4 > 3) Linv[j0 + 3] = i3;
        }
        fk_sync();
    }

    // ---------------------------------------------------------------------------------------------
    // one stage of the backward Riccati sweep with factorisation (OCP_QP_FACT_SOLVE_KKT_STEP, x_ocp_qp_kkt.c:880-966),
    // right-hand side = residual set 0.  ML holds L_{k+1} (row n1 = its gradient row) on entry and L_k on exit.
    //   [A; b'] -> MA, in place  AL = [A; b'] * Lxx_{k+1}                              (TRMM_RLNN)
    //   8-column tiles:  acc = H + diag + AL AL' - (columns already factorised)         (SYRK + left-looking POTRF)
    // ---------------------------------------------------------------------------------------------
    template <int KIND>
    FK_DEV void fact_stage(int k)
    {
        constexpr int nx = KD<KIND>::nx, nu = KD<KIND>::nu, n = nx + nu, nx1 = KD<KIND>::nx1;
        constexpr int RP = (n + 1 + G - 1) / G;
        const StageDesc &sd = sdk<KIND>();
        const View v = view<KIND>(k);
        const int nb = sd.nb, ns = sd.ns, nc = sd.nc;
        const int *idxb = v.ip + sd.idx_off, *rev = idxb + nb;
        const int nu1 = (nx1 > 0 && k + 1 < A.N) ? NU : 0, n1 = nx1 + nu1;
        double *Gam = V, *gam = Gam + A.nce, *tmp0 = gam + A.nce, *tmp1 = tmp0 + A.nbe;
        double *dadd = tmp1 + A.nbe, *rowv = dadd + evn(NM + 1), *Linv = rowv + evn(NM + 1);
        double *Zi = Linv + evn(NM + 1), *ds = Zi + A.ns2e, *lnx = ds + A.ns2e;
        // ---- stage inputs: [A; b'] into MA (asynchronous), gradient, constraint quantities
        if (nx1 > 0)
        {
            stage_mat<n, nx1>(MA, v.q + sd.q_BAt);
            const double *b_ = v.w + sd.res.b;
            for (int j = li; j < nx1; j += G) fk_cp8(MA + n + LD * j, b_ + j);
            // gradient row of L_{k+1} (x part) before ML is overwritten
            for (int j = li; j < nx1; j += G) lnx[j] = ML[n1 + LD * (nu1 + j)];
        }
        {
            const double *g0 = v.w + sd.res.g;
            for (int i = li; i < n; i += G) { rowv[i] = g0[i]; dadd[i] = A.o.reg_prim; }
        }
        {
            // Gamma, gamma (COMPUTE_GAMMA_GAMMA_QP, x_core_qp_ipm_aux.c:38-86)
            const double *gl = v.s + sd.sol.lam, *gt = v.s + sd.sol.t, *grd = v.w + sd.res.d, *grm = v.w + sd.res.m;
            const double t_min_inv = A.o.t_min > 0 ? 1.0 / A.o.t_min : 1e30;
            for (int i = li; i < nc; i += G)
            {
                const double l = gl[i], tt = gt[i], ti = 1.0 / tt;
                if (A.o.t_lam_min == 1)
                    Gam[i] = (tt < A.o.t_min ? t_min_inv : ti) * (l < A.o.lam_min ? A.o.lam_min : l);
                else
                    Gam[i] = ti * l;
                gam[i] = ti * (grm[i] - l * grd[i]);
            }
        }
        fk_sync();
        if (ns > 0)
        {
            cond_slacks(nb, ns, rev, v.q + sd.q_Z, 1, Gam, gam, v.w + sd.res.g + n, Zi, ds, tmp0, tmp1);
            fk_sync();
            for (int j = li; j < 2 * ns; j += G)
            {
                st(v.w + sd.w_Zsi + j, Zi[j]);
                st(v.w + sd.step.ux + n + j, ds[j]);
            }
        }
        else
        {
            for (int i = li; i < nb; i += G)
            {
                tmp0[i] = Gam[i] + Gam[nb + i];
                tmp1[i] = gam[i] - gam[nb + i];
            }
            fk_sync();
        }
        for (int i = li; i < nb; i += G)
        {
            const int ix = idxb[i];
            dadd[ix] += tmp0[i];
            rowv[ix] += tmp1[i];
        }
        fk_cp_wait();
        fk_sync();
        if (nx1 > 0)
        {
            // ---- in place: AL = [A; b'] * Lxx   (row slots x 8-column tiles; Lxx(c, j) = Lx[c + LD*j], lower triangular)
            const double *Lx = ML + nu1 + LD * nu1;
#pragma unroll
            for (int jt = 0; jt < nx1; jt += 8)
            {
                const int w = nx1 - jt < 8 ? nx1 - jt : 8;
                double acc[RPM][8];
#pragma unroll
                for (int m = 0; m < RP; m++)
#pragma unroll
                    for (int q = 0; q < 8; q++) acc[m][q] = 0.0;
                // triangular head: column jt+q of Lxx starts at row jt+q
#pragma unroll
                for (int h = 0; h < 8; h++)
                {
                    if (h >= w) continue;
                    const int c = jt + h;
                    double a[RPM];
#pragma unroll
                    for (int m = 0; m < RP; m++) a[m] = MA[li + G * m + LD * c];
#pragma unroll
                    for (int q = 0; q <= h; q++)
                    {
                        const double l = Lx[c + LD * (jt + q)];
#pragma unroll
                        for (int m = 0; m < RP; m++) acc[m][q] += a[m] * l;
                    }
                }
#pragma unroll 2
                for (int c = jt + 8; c < nx1; c++)
                {
                    double a[RPM];
#pragma unroll
                    for (int m = 0; m < RP; m++) a[m] = MA[li + G * m + LD * c];
#pragma unroll
                    for (int q = 0; q < 8; q++)
                    {
                        const double l = Lx[c + LD * (jt + q)];
#pragma unroll
                        for (int m = 0; m < RP; m++) acc[m][q] += a[m] * l;
                    }
                }
#pragma unroll
                for (int m = 0; m < RP; m++)
                {
                    const int r = li + G * m;
                    if (r <= n)
                    {
#pragma unroll
                        for (int q = 0; q < 8; q++)
                            if (q < w) MA[r + LD * (jt + q)] = acc[m][q];
                    }
                }
            }
            fk_sync();
            // Pb = Lxx * (Lxx' b)  (row n of AL is Lxx' b at this point), then the gradient row gets l_{k+1}
            {
                double *Pb = v.w + sd.w_Pb;
                for (int i = li; i < nx1; i += G)
                {
                    double s0 = 0.0, s1 = 0.0;
                    int c = 0;
                    for (; c + 1 <= i; c += 2) { s0 += Lx[i + LD * c] * MA[n + LD * c]; s1 += Lx[i + LD * (c + 1)] * MA[n + LD * (c + 1)]; }
                    if (c <= i) s0 += Lx[i + LD * c] * MA[n + LD * c];
                    st(Pb + i, s0 + s1);
                }
            }
            fk_sync();
            for (int j = li; j < nx1; j += G) MA[n + LD * j] += lnx[j];
            fk_sync();
        }
        // ---- column tiles: SYRK + left-looking Cholesky, rows r >= jt (row n = gradient row)
        const double *Hg = v.q + sd.q_RSQ;
        double *Lg = v.w + sd.w_L, *lrow = v.w + sd.w_lrow;
#pragma unroll
        for (int jt = 0; jt < n; jt += 8)
        {
            const int w = n - jt < 8 ? n - jt : 8;
            const int m0 = jt / G;                      // first row slot that reaches into the tile
            double acc[RPM][8], h[RPM][8];
            // H (lower, from global; issued first so that the loads overlap the products), diagonal additions, gradient
#pragma unroll
            for (int m = 0; m < RP; m++)
            {
                if (m < m0) continue;
                const int r = li + G * m;
#pragma unroll
                for (int q = 0; q < 8; q++)
                {
                    acc[m][q] = 0.0;
                    double hv = 0.0;
                    if (q < w)
                    {
                        if (r < n) { if (r >= jt + q) hv = fk_ldg(Hg + r + n * (jt + q)); }
                        else if (r == n) hv = rowv[jt + q];
                        if (r == jt + q) hv += dadd[r];
                    }
                    h[m][q] = hv;
                }
            }
            if (nx1 > 0)
            {
#pragma unroll 3
                for (int c = 0; c < nx1; c++)
                {
                    double a[RPM], b[8];
#pragma unroll
                    for (int m = 0; m < RP; m++) a[m] = m >= m0 ? MA[li + G * m + LD * c] : 0.0;
#pragma unroll
                    for (int q = 0; q < 8; q++) b[q] = q < w ? MA[jt + q + LD * c] : 0.0;
#pragma unroll
                    for (int m = 0; m < RP; m++)
                        if (m >= m0)
#pragma unroll
                            for (int q = 0; q < 8; q++) acc[m][q] += a[m] * b[q];
                }
            }
#pragma unroll 2
            for (int c = 0; c < jt; c++)
            {
                double a[RPM], b[8];
#pragma unroll
                for (int m = 0; m < RP; m++) a[m] = m >= m0 ? ML[li + G * m + LD * c] : 0.0;
#pragma unroll
                for (int q = 0; q < 8; q++) b[q] = q < w ? ML[jt + q + LD * c] : 0.0;
#pragma unroll
                for (int m = 0; m < RP; m++)
                    if (m >= m0)
#pragma unroll
                        for (int q = 0; q < 8; q++) acc[m][q] -= a[m] * b[q];
            }
#pragma unroll
            for (int m = 0; m < RP; m++)
                if (m >= m0)
#pragma unroll
                    for (int q = 0; q < 8; q++) acc[m][q] += h[m][q];
            // ---- first half of the tile
            if (w >= 4) panel4<n, RP, 4>(jt, m0, acc, 0, Lg, lrow, Linv);
            else if (w == 3) panel4<n, RP, 3>(jt, m0, acc, 0, Lg, lrow, Linv);
            else if (w == 2) panel4<n, RP, 2>(jt, m0, acc, 0, Lg, lrow, Linv);
            else panel4<n, RP, 1>(jt, m0, acc, 0, Lg, lrow, Linv);
            if (w > 4)
            {
                // ---- second half: update with the four columns just finished, then its own panel
                const int m1 = (jt + 4) / G;
#pragma unroll
                for (int c = 0; c < 4; c++)
                {
                    double b[4];
#pragma unroll
                    for (int q = 0; q < 4; q++) b[q] = 4 + q < w ? ML[jt + 4 + q + LD * (jt + c)] : 0.0;
#pragma unroll
                    for (int m = 0; m < RP; m++)
                        if (m >= m1)
#pragma unroll
                            for (int q = 0; q < 4; q++) acc[m][4 + q] -= acc[m][c] * b[q];
                }
                if (w >= 8) panel4<n, RP, 4>(jt + 4, m1, acc, 4, Lg, lrow, Linv);
                else if (w == 7) panel4<n, RP, 3>(jt + 4, m1, acc, 4, Lg, lrow, Linv);
                else if (w == 6) panel4<n, RP, 2>(jt + 4, m1, acc, 4, Lg, lrow, Linv);
                else panel4<n, RP, 1>(jt + 4, m1, acc, 4, Lg, lrow, Linv);
            }
        }
        {
            double *li_ = v.w + sd.w_Linv;
            for (int j = li; j < n; j += G) st(li_ + j, Linv[j]);
        }
        fk_sync();
    }

    FK_DEV void fact_backward()
    {
        fk_sync();
        fact_stage<2>(A.N);
        for (int k = A.N - 1; k >= 1; k--) fact_stage<1>(k);
        fact_stage<0>(0);
    }

    // ---------------------------------------------------------------------------------------------
    // one stage of the backward substitution with the existing factorisation (OCP_QP_SOLVE_KKT_STEP, x_ocp_qp_kkt.c:
    // 1582-1680), right-hand side = residual set 0, result (backward quantities) into the step.  The complementarity
    // right-hand side of the corrector (x_core_qp_ipm_aux.c:695-754) is formed on the fly:
    // rm_mode 1: res_m = bkp + dt*dlam - sigma_mu; 2: res_m = bkp - sigma_mu.
    // ---------------------------------------------------------------------------------------------
    template <int KIND>
    FK_DEV void solve_stage(int k, int rm_mode, double sigma_mu, bool stw)
    {
        constexpr int nx = KD<KIND>::nx, nu = KD<KIND>::nu, n = nx + nu, nx1 = KD<KIND>::nx1;
        constexpr int nsolve = nu;                  // stage 0 has nx = 0: n = nu
        const StageDesc &sd = sdk<KIND>();
        const View v = view<KIND>(k);
        const int nb = sd.nb, ns = sd.ns, nc = sd.nc;
        const int *idxb = v.ip + sd.idx_off, *rev = idxb + nb;
        double *vv = V, *gam = vv + A.nve, *Gam = gam + A.nce, *tmp0 = Gam + A.nce, *tmp1 = tmp0 + A.nbe;
        double *Zi = tmp1 + A.nbe, *ds = Zi + A.ns2e, *xprev = ds + A.ns2e, *tmpx = xprev + evn(NX);
        double *Lis = tmpx + evn(NX), *pbs = Lis + evn(NM + 1);
        const bool so = act && stw;
        if (nx1 > 0) stage_mat<n, nx1>(MA, v.q + sd.q_BAt);
        {
            const double *Lg = v.w + sd.w_L;
            for (int e = li; e < n * nsolve; e += G)
            {
                const int c = e / (n > 0 ? n : 1), r = e - c * n;
                fk_cp8(LU + r + LD * c, Lg + e);
            }
            const double *g0 = v.w + sd.res.g, *Li = v.w + sd.w_Linv, *zs = v.w + sd.w_Zsi, *pb = v.w + sd.w_Pb;
            for (int i = li; i < n; i += G) vv[i] = g0[i];
            for (int i = li; i < nsolve; i += G) Lis[i] = Li[i];
            for (int j = li; j < 2 * ns; j += G) Zi[j] = zs[j];
            for (int j = li; j < nx1; j += G) pbs[j] = pb[j];
            const double *gl = v.s + sd.sol.lam, *gt = v.s + sd.sol.t, *grd = v.w + sd.res.d, *gm = v.q + sd.q_dmask;
            double *grm = v.w + sd.res.m;
            const double *bk = v.w + sd.w_rmb, *dl = v.w + sd.step.lam, *dtt = v.w + sd.step.t;
            const double t_min_inv = A.o.t_min > 0 ? 1.0 / A.o.t_min : 1e30;
            for (int i = li; i < nc; i += G)
            {
                const double l = gl[i], tt = gt[i], ti = 1.0 / tt;
                double m = rm_mode == 1 ? bk[i] + dtt[i] * dl[i] - sigma_mu : bk[i] - sigma_mu;
                m *= fk_ldg(gm + i);
                if (so) grm[i] = m;
                // the slack elimination needs the Gamma of the factorisation (clipped when t_lam_min==1)
                Gam[i] = (ns > 0 && A.o.t_lam_min == 1) ? (tt < A.o.t_min ? t_min_inv : ti) * (l < A.o.lam_min ? A.o.lam_min : l) : ti * l;
                gam[i] = ti * (m - l * grd[i]);
            }
        }
        fk_sync();
        if (ns > 0)
        {
            cond_slacks(nb, ns, rev, v.q + sd.q_Z, 0, Gam, gam, v.w + sd.res.g + n, Zi, ds, tmp0, tmp1);
            fk_sync();
            double *o_ = v.w + sd.step.ux + n;
            for (int j = li; j < 2 * ns; j += G)
                if (so) o_[j] = ds[j];
        }
        else
        {
            for (int i = li; i < nb; i += G) tmp1[i] = gam[i] - gam[nb + i];
            fk_sync();
        }
        for (int i = li; i < nb; i += G) vv[idxb[i]] += tmp1[i];
        for (int j = li; j < nx1; j += G) tmpx[j] = xprev[j] + pbs[j];
        fk_cp_wait();
        fk_sync();
        if (nx1 > 0)
        {
            for (int i = li; i < n; i += G)
            {
                double s0 = 0.0, s1 = 0.0;
                const double *arow = MA + i;
                int j = 0;
#pragma unroll 4
                for (; j + 1 < nx1; j += 2) { s0 += arow[LD * j] * tmpx[j]; s1 += arow[LD * (j + 1)] * tmpx[j + 1]; }
                if (j < nx1) s0 += arow[LD * j] * tmpx[j];
                vv[i] += s0 + s1;
            }
            fk_sync();
        }
        // TRSV_LNN(_MN): forward substitution on the first nsolve unknowns (redundantly by every lane), then the rows below
        {
            double u[NU > 0 ? NU : 1];
#pragma unroll
            for (int j = 0; j < nsolve; j++)
            {
                double part = 0.0;
#pragma unroll
                for (int c = 0; c < j; c++) part += LU[j + LD * c] * u[c];
                u[j] = (vv[j] - part) * Lis[j];
            }
            fk_sync();
            for (int i = li; i < n; i += G)
            {
                double x = vv[i];
                if (i < nsolve)
                {
#pragma unroll
                    for (int j = 0; j < nsolve; j++)
                        if (i == j) x = u[j];
                }
                else
                {
                    double part = 0.0;
#pragma unroll
                    for (int c = 0; c < nsolve; c++) part += LU[i + LD * c] * u[c];
                    x -= part;
                }
                vv[i] = x;
                if (so) (v.w + sd.step.ux)[i] = x;
                if (i >= nu) xprev[i - nu] = x;
            }
        }
        fk_sync();
    }

    FK_DEV void solve_backward(int rm_mode, double sigma_mu, bool stw)
    {
        fk_sync();
        solve_stage<2>(A.N, rm_mode, sigma_mu, stw);
        for (int k = A.N - 1; k >= 1; k--) solve_stage<1>(k, rm_mode, sigma_mu, stw);
        solve_stage<0>(0, rm_mode, sigma_mu, stw);
    }

    // ---------------------------------------------------------------------------------------------
    // one stage of the forward sweep (x_ocp_qp_kkt.c:968-1006 / 1682-1722) + step of the constraint variables
    // (:1176-1193, EXPAND_SLACKS :524-598, COMPUTE_LAM_T_QP x_core_qp_ipm_aux.c:164-189) + the ratio test
    // (COMPUTE_ALPHA_QP :375-398) + the residual of the linear system (OCP_QP_RES_COMPUTE_LIN) -> residual set 1.
    // after_fact: start from -lrow, pi = P x + p with p from lrow; else: start from the backward quantities stored in
    // the step, pi = p_backward + P x.
    // ---------------------------------------------------------------------------------------------
    struct FwdAcc
    {
        double alpha, m0, m1, m2, m3;
        int f0, f1, f2, f3;
    };

    template <int KIND>
    FK_DEV void fwd_stage(int k, int after_fact, int do_lin, bool stw, FwdAcc &F)
    {
        constexpr int nx = KD<KIND>::nx, nu = KD<KIND>::nu, n = nx + nu, nx1 = KD<KIND>::nx1;
        constexpr int nsolve = nu;
        const StageDesc &sd = sdk<KIND>();
        const View v = view<KIND>(k);
        const int nb = sd.nb, ns = sd.ns, nc = sd.nc;
        const int *idxb = v.ip + sd.idx_off, *rev = idxb + nb;
        const int nu1 = (nx1 > 0 && k + 1 < A.N) ? NU : 0;
        const int NXe = evn(NX);
        double *vv = V, *x1 = vv + A.nve, *tmp = x1 + NXe, *p1 = tmp + NXe, *pik = p1 + NXe, *pim = pik + NXe;
        double *Gam = pim + NXe, *dt = Gam + A.nce, *lam = dt + A.nce, *dlm = lam + A.nce;
        double *Zi = dlm + A.nce, *ds = Zi + A.ns2e, *g_ = ds + A.ns2e, *tmp0 = g_ + A.nve;
        double *Lis = tmp0 + A.nbe, *bs = Lis + evn(NM + 1), *ts = bs + NXe, *rds = ts + A.nce, *rms = rds + A.nce, *mks = rms + A.nce;
        const bool so = act && stw;
        // ---- staging: BAt, first nsolve columns of L_k, L_{k+1}, vectors
        if (nx1 > 0)
        {
            stage_mat<n, nx1>(MA, v.q + sd.q_BAt);
            const StageDesc &s1 = sdr(k + 1);
            const View v1 = viewr(k + 1);
            const double *L1 = v1.w + s1.w_L;
            if (nu1 > 0) stage_mat<NX + NU, NX + NU>(ML, L1);
            else stage_mat<NX, NX>(ML, L1);
            const double *ps = after_fact ? v1.w + s1.w_lrow + nu1 : v1.w + s1.step.ux + nu1;   // p part / backward value of x_{k+1}
            for (int j = li; j < nx1; j += G) p1[j] = ps[j];
            const double *b_ = v.w + sd.res.b;
            for (int j = li; j < nx1; j += G) bs[j] = b_[j];
        }
        {
            const double *Lg = v.w + sd.w_L;
            for (int e = li; e < n * nsolve; e += G)
            {
                const int c = e / (n > 0 ? n : 1), r = e - c * n;
                fk_cp8(LU + r + LD * c, Lg + e);
            }
            const double *src = after_fact ? v.w + sd.w_lrow : v.w + sd.step.ux, *Li = v.w + sd.w_Linv;
            for (int i = li; i < nsolve; i += G) { vv[i] = -src[i]; Lis[i] = Li[i]; }
            // x part (k>0) was written into vv by the previous stage
            if (ns > 0)
            {
                const double *z_ = v.w + sd.w_Zsi, *d_ = v.w + sd.step.ux + n;
                for (int j = li; j < 2 * ns; j += G) { Zi[j] = z_[j]; ds[j] = d_[j]; }
            }
            const double *gl = v.s + sd.sol.lam, *gt = v.s + sd.sol.t, *grd = v.w + sd.res.d, *grm = v.w + sd.res.m, *gm = v.q + sd.q_dmask;
            for (int i = li; i < nc; i += G)
            {
                lam[i] = gl[i];
                ts[i] = gt[i];
                rds[i] = grd[i];
                rms[i] = grm[i];
                mks[i] = fk_ldg(gm + i);
            }
        }
        fk_cp_wait();
        fk_sync();
        // ---- TRSV_LTN(_MN): u = -Luu^{-T} (l_u + Lxu' x): the dot products over the x rows by the group, the small triangle
        // redundantly by every lane
        if (nsolve > 0)
        {
            double wv[NU > 0 ? NU : 1];
#pragma unroll
            for (int j = 0; j < nsolve; j++)
            {
                double part = 0.0;
                for (int i = nsolve + li; i < n; i += G) part += LU[i + LD * j] * vv[i];
                if (n > nsolve) part = gsum(part);
                wv[j] = vv[j] - part;
            }
#pragma unroll
            for (int j = nsolve - 1; j >= 0; j--)
            {
                double part = 0.0;
#pragma unroll
                for (int i = j + 1; i < nsolve; i++) part += LU[i + LD * j] * wv[i];
                wv[j] = (wv[j] - part) * Lis[j];
            }
            fk_sync();
#pragma unroll
            for (int j = 0; j < nsolve; j++)
                if (li == j % G) vv[j] = wv[j];
            fk_sync();
        }
        {
            double *o_ = v.w + sd.step.ux;
            for (int i = li; i < n; i += G)
                if (so) o_[i] = vv[i];
        }
        if (nx1 > 0)
        {
            const double *Lx = ML + nu1 + LD * nu1;                          // Lxx of stage k+1
            double *ob = v.w + sd.ires.b;
            for (int j = li; j < nx1; j += G)
            {
                double s0 = 0.0, s1 = 0.0;
                const double *acol = MA + LD * j;
                int i = 0;
#pragma unroll 4
                for (; i + 1 < n; i += 2) { s0 += acol[i] * vv[i]; s1 += acol[i + 1] * vv[i + 1]; }
                if (i < n) s0 += acol[i] * vv[i];
                const double acc = s0 + s1, bv = bs[j];
                const double xj = bv + acc;
                x1[j] = xj;
                if (do_lin)
                {
                    const double r = bv - xj + acc;
                    if (so) ob[j] = r;
                    const double a = fabs(r);
                    F.m1 = fmax(F.m1, a);
                    F.f1 |= (a != a);
                }
            }
            fk_sync();
            for (int j = li; j < nx1; j += G)
            {
                double s0 = 0.0, s1 = 0.0;
                const double *lcol = Lx + LD * j;
                int i = j;
                for (; i + 1 < nx1; i += 2) { s0 += lcol[i] * x1[i]; s1 += lcol[i + 1] * x1[i + 1]; }
                if (i < nx1) s0 += lcol[i] * x1[i];
                const double acc = s0 + s1;
                tmp[j] = after_fact ? acc + p1[j] : acc;
            }
            fk_sync();
            double *pi = v.w + sd.step.pi;
            for (int i = li; i < nx1; i += G)
            {
                double s0 = 0.0, s1 = 0.0;
                const double *lrow_ = Lx + i;
                int c = 0;
                for (; c + 1 <= i; c += 2) { s0 += lrow_[LD * c] * tmp[c]; s1 += lrow_[LD * (c + 1)] * tmp[c + 1]; }
                if (c <= i) s0 += lrow_[LD * c] * tmp[c];
                const double acc = s0 + s1;
                const double pv = after_fact ? acc : acc + p1[i];
                if (so) pi[i] = pv;
                pik[i] = pv;
            }
        }
        // ---- constraint part of the step at this stage
        {
            const double t_min_inv = A.o.t_min > 0 ? 1.0 / A.o.t_min : 1e30;
            for (int i = li; i < nc; i += G)
            {
                const double l = lam[i], tt = ts[i];
                Gam[i] = (ns > 0 && A.o.t_lam_min == 1) ? (tt < A.o.t_min ? t_min_inv : 1.0 / tt) * (l < A.o.lam_min ? A.o.lam_min : l)
                                                      : (1.0 / tt) * l;
            }
            for (int i = li; i < nb; i += G)
            {
                const double a = vv[idxb[i]];
                dt[i] = a;
                dt[nb + i] = -a;
            }
            if (ns > 0)
            {
                fk_sync();
                for (int j = li; j < 2 * ns; j += G)
                {
                    const int jj = j < ns ? j : j - ns, offc = j < ns ? 0 : nb;
                    double d = ds[j];
                    for (int i = 0; i < nb; i++)
                        if (rev[i] == jj) d += Gam[offc + i] * dt[offc + i];
                    d = -Zi[j] * d;
                    ds[j] = d;
                    dt[2 * nb + j] = d;
                }
                fk_sync();
                for (int i = li; i < 2 * nb; i += G)
                {
                    const int up = i >= nb, ii = up ? i - nb : i;
                    if (rev[ii] >= 0) dt[i] += ds[(up ? ns : 0) + rev[ii]];
                }
                double *o_ = v.w + sd.step.ux + n;
                for (int j = li; j < 2 * ns; j += G)
                    if (so) o_[j] = ds[j];
            }
            fk_sync();
            double *odl = v.w + sd.step.lam, *odt = v.w + sd.step.t, *ld_ = v.w + sd.ires.d, *lm_ = v.w + sd.ires.m;
            for (int i = li; i < nc; i += G)
            {
                const double l = lam[i], tt = ts[i], ti = 1.0 / tt, rdi = rds[i], rmi = rms[i];
                const double dtr = dt[i];
                double dl = -ti * (rmi + (l * dtr) - (l * rdi));
                double dti = dtr - rdi;
                const double mk = mks[i];
                dl *= mk;
                dti *= mk;
                if (so) { odl[i] = dl; odt[i] = dti; }
                dlm[i] = dl * mk;     // masked step multipliers (tmp_lam_mask of the linear residual)
                // ratio test (min over constraints, see COMPUTE_ALPHA_QP)
                if (l + dl < 0.0) F.alpha = fmin(F.alpha, -l / dl);
                if (tt + dti < 0.0) F.alpha = fmin(F.alpha, -tt / dti);
                if (do_lin)
                {
                    // res_d = rhs_d + dt -/+ (v[idxb] | C'v) [- ds] = rhs_d + dt - dtr ;  res_m = rhs_m + lam dt + dlam t
                    double r = (dti + rdi) - dtr;
                    r *= mk;
                    if (so) ld_[i] = r;
                    double a = fabs(r);
                    F.m2 = fmax(F.m2, a);
                    F.f2 |= (a != a);
                    double mm = rmi + l * dti + dl * tt;
                    mm *= mk;
                    if (so) lm_[i] = mm;
                    a = fabs(mm);
                    F.m3 = fmax(F.m3, a);
                    F.f3 |= (a != a);
                }
            }
        }
        fk_sync();
        if (do_lin)
        {
            // ---- res_g of the linear system (lane = row): H dux + rhs_g - dpi_{k-1} + A dpi_k + constraint multipliers
            const double *Hg = v.q + sd.q_RSQ, *gv = v.w + sd.res.g;
            for (int i = li; i < nb; i += G) tmp0[i] = dlm[nb + i] - dlm[i];
            fk_sync();
            for (int i = li; i < n; i += G)
            {
                double r = gdot_sym<n>(Hg, i, vv) + gv[i];
                if (nx > 0 && i >= nu) r -= pim[i - nu];
                if (nx1 > 0)
                {
                    double s0 = 0.0, s1 = 0.0;
                    const double *arow = MA + i;
                    int j = 0;
#pragma unroll 4
                    for (; j + 1 < nx1; j += 2) { s0 += arow[LD * j] * pik[j]; s1 += arow[LD * (j + 1)] * pik[j + 1]; }
                    if (j < nx1) s0 += arow[LD * j] * pik[j];
                    r += s0 + s1;
                }
                g_[i] = r;
            }
            fk_sync();
            for (int i = li; i < nb; i += G) g_[idxb[i]] += tmp0[i];
            if (ns > 0)
            {
                const double *Z = v.q + sd.q_Z, *zv = v.w + sd.res.g + n;
                for (int j = li; j < 2 * ns; j += G)
                {
                    double r = fk_ldg(Z + j) * ds[j] + zv[j] - dlm[2 * nb + j];
                    const int jj = j < ns ? j : j - ns, offl = j < ns ? 0 : nb;
                    for (int i = 0; i < nb; i++)
                        if (rev[i] == jj) r -= dlm[offl + i];
                    g_[n + j] = r;
                }
            }
            fk_sync();
            double *og = v.w + sd.ires.g;
            for (int i = li; i < n + 2 * ns; i += G)
            {
                const double r = g_[i];
                if (so) og[i] = r;
                const double a = fabs(r);
                F.m0 = fmax(F.m0, a);
                F.f0 |= (a != a);
            }
        }
        if (nx1 > 0)
        {
            for (int j = li; j < nx1; j += G)
            {
                vv[nu1 + j] = x1[j];
                pim[j] = pik[j];
            }
        }
        fk_sync();
    }

    // returns the step length; lin_nrm = inf-norms of the residual of the linear system (do_lin)
    FK_DEV double forward_pass(int after_fact, int do_lin, bool stw, double lin_nrm[4])
    {
        FwdAcc F;
        F.alpha = 1.0;
        F.m0 = F.m1 = F.m2 = F.m3 = 0.0;
        F.f0 = F.f1 = F.f2 = F.f3 = 0;
        fk_sync();
        fwd_stage<0>(0, after_fact, do_lin, stw, F);
        for (int k = 1; k < A.N; k++) fwd_stage<1>(k, after_fact, do_lin, stw, F);
        fwd_stage<2>(A.N, after_fact, do_lin, stw, F);
        if (do_lin)
        {
            lin_nrm[0] = gmax_nan(F.m0, F.f0);
            lin_nrm[1] = gmax_nan(F.m1, F.f1);
            lin_nrm[2] = gmax_nan(F.m2, F.f2);
            lin_nrm[3] = gmax_nan(F.m3, F.f3);
        }
        return gmin(F.alpha);
    }

    // COMPUTE_MU_AFF_QP (x_core_qp_ipm_aux.c:636-668)
    FK_DEV double mu_aff_pass(double alpha)
    {
        double acc = 0.0;
        for (int k = 0; k <= A.N; k++)
        {
            const StageDesc &s = sdr(k);
            const View v = viewr(k);
            const double *l = v.s + s.sol.lam, *t = v.s + s.sol.t, *dl = v.w + s.step.lam, *dt = v.w + s.step.t;
            for (int i = li; i < s.nc; i += G) acc += fabs((l[i] + alpha * dl[i]) * (t[i] + alpha * dt[i]));
        }
        return gsum(acc) * nc_mask_inv;
    }

    // OCP_QP_INIT_VAR, var_init_scheme 1 (x_ocp_qp_ipm.c:1611-1760,1884-2022); no general constraints here
    FK_DEV void init_var()
    {
        const double thr0 = 0.1;
        const int N = A.N;
        // the reference's plugin zeroes the primal iterate before every solve, whatever warm_start says
        // (acados/ocp_qp/ocp_qp_hpipm.c:333-336): warm starts carry over pi, lam and t only
        if (A.o.warm_start >= 2)
        {
            const double lmin = A.o.warm_start >= 3 ? A.o.lam0_min : thr0, tmin = A.o.warm_start >= 3 ? A.o.t0_min : thr0;
            for (int k = 0; k <= N; k++)
            {
                const StageDesc &s = sdr(k);
                const View v = viewr(k);
                double *l = v.s + s.sol.lam, *t = v.s + s.sol.t, *gux = v.s + s.sol.ux, *gpi = v.s + s.sol.pi;
                // keep what the caller passed in for the case that this QP is handed back to the generic kernel
                double *kl = v.w + s.itref.lam, *kt = v.w + s.itref.t, *kp = v.w + s.itref.pi;
                for (int i = li; i < s.n + 2 * s.ns; i += G) st(gux + i, 0.0);
                for (int i = li; i < s.nx1; i += G) st(kp + i, gpi[i]);
                for (int i = li; i < s.nc; i += G)
                {
                    st(kl + i, l[i]);
                    st(kt + i, t[i]);
                    if (l[i] < lmin) st(l + i, lmin);
                    if (t[i] < tmin) st(t + i, tmin);
                }
            }
            fk_sync();
            return;
        }
        double *ux = V, *tt = ux + A.nve;
        for (int k = 0; k <= N; k++)
        {
            const StageDesc &s = sdr(k);
            const View v = viewr(k);
            const int n = s.n, nb = s.nb, ns = s.ns, nc = s.nc;
            const int *idxb = v.ip + s.idx_off, *rev = idxb + nb;
            const double *d = v.q + s.q_d;
            double *gux = v.s + s.sol.ux, *gpi = v.s + s.sol.pi, *gl = v.s + s.sol.lam, *gt = v.s + s.sol.t;
            for (int i = li; i < s.nx1; i += G) st(gpi + i, 0.0);
            if (A.o.t0_init == 0 || A.o.t0_init == 1)
            {
                const double l0 = A.o.t0_init == 0 ? sqrt(A.o.mu0) : A.o.mu0, t0 = A.o.t0_init == 0 ? sqrt(A.o.mu0) : 1.0;
                for (int i = li; i < n + 2 * ns; i += G) st(gux + i, 0.0);
                for (int i = li; i < nc; i += G) { st(gl + i, l0); st(gt + i, t0); }
                continue;
            }
            for (int i = li; i < n + 2 * ns; i += G) ux[i] = 0.0;
            fk_sync();
            for (int j = li; j < 2 * ns; j += G)
            {
                double tj = ux[n + j] - d[2 * nb + j];
                if (tj < thr0)
                {
                    tj = thr0;
                    ux[n + j] = d[2 * nb + j] + tj;
                }
                tt[2 * nb + j] = tj;
            }
            fk_sync();
            for (int j = li; j < nb; j += G)
            {
                const int ix = idxb[j];
                double tl = ux[ix], tu = -ux[ix];
                if (ns > 0 && rev[j] != -1) { tl += ux[n + rev[j]]; tu += ux[n + ns + rev[j]]; }
                tl -= d[j];
                tu -= d[nb + j];
                if (tl < thr0)
                {
                    if (tu < thr0)
                    {
                        ux[ix] = 0.5 * (d[j] - d[nb + j]);
                        tl = thr0; tu = thr0;
                    }
                    else
                    {
                        tl = thr0;
                        ux[ix] = d[j] + thr0;
                    }
                }
                else if (tu < thr0)
                {
                    tu = thr0;
                    ux[ix] = -d[nb + j] - thr0;
                }
                tt[j] = tl;
                tt[nb + j] = tu;
            }
            fk_sync();
            for (int i = li; i < n + 2 * ns; i += G) st(gux + i, ux[i]);
            for (int i = li; i < nc; i += G)
            {
                st(gt + i, tt[i]);
                st(gl + i, A.o.mu0 / tt[i]);
            }
            fk_sync();
        }
        fk_sync();
    }

    double nc_mask_inv;

    // ---------------------------------------------------------------------------------------------
    // driver (OCP_QP_IPM_SOLVE x_ocp_qp_ipm.c:2684-3120 + OCP_QP_IPM_DELTA_STEP :2208-2682) for the QPs of this warp;
    // q = index of this group's QP (clamped to a valid one; valid = it exists)
    // ---------------------------------------------------------------------------------------------
    FK_DEV void solve(int q, bool valid)
    {
        const int N = A.N;
        const int SM = CUIPM_STAT_M;
        qp = A.qp + (size_t) q * A.qp_stride;
        sol = A.sol + (size_t) q * A.sol_stride;
        wk = A.work + (size_t) q * A.work_stride;
        act = valid;
        cuipm_info *info = A.info + q;
        double *stat = A.stat ? A.stat + (size_t) q * SM * (A.o.stat_max + 1) : nullptr;
        QpState Q;
        Q.mu = Q.obj = Q.gap = 0.0; Q.alpha = 1.0; Q.res_m_tau = 0.0;
        Q.res_max[0] = Q.res_max[1] = Q.res_max[2] = Q.res_max[3] = 0.0;
        if (stat && act)
            for (int i = li; i < SM * (A.o.stat_max + 1); i += G) stat[i] = 0.0;

        // constraint mask census (x_ocp_qp_ipm.c:2774-2806)
        int cnt = 0;
        for (int k = 0; k <= N; k++)
        {
            const StageDesc &s = sdr(k);
            const double *gm = viewr(k).q + s.q_dmask;
            for (int i = li; i < s.nc; i += G) cnt += fk_ldg(gm + i) != 0.0;
        }
        const int nc_mask = (int) (gsum((double) cnt) + 0.5);
        nc_mask_inv = nc_mask > 0 ? 1.0 / nc_mask : 0.0;
        bool redo = nc_mask == 0;          // no active constraint: the unconstrained branch lives in the generic kernel

        if (act && redo) { hand_back(q, info); act = false; }
        init_var();
        // masked constraints start with zero multipliers (the generic kernel does this only when some constraint is masked;
        // multiplying by 1.0 is exact)
        for (int k = 0; k <= N; k++)
        {
            const StageDesc &s = sdr(k);
            const View v = viewr(k);
            double *l = v.s + s.sol.lam;
            const double *gm = v.q + s.q_dmask;
            for (int i = li; i < s.nc; i += G) st(l + i, l[i] * fk_ldg(gm + i));
        }
        fk_sync();
        // Every sweep has exactly one call site (everything is inlined into the kernel, and the hot code should exist once):
        // the residual sweep opens each pass of the loop (pass 0: residuals of the initial point; pass kk: move along the
        // step of iteration kk-1, then residuals), the predictor / corrector / conditional corrector are phases 0 / 1 / 2 of
        // one inner loop.
        for (int kk = 0;; kk++)
        {
            res_pass(kk > 0, Q.alpha, Q);
            if (stat && act && kk < A.o.stat_max && li == 0)
            {
                double *sr = stat + SM * (size_t) kk;
                if (kk > 0) sr[6] = Q.mu;
                sr[7] = Q.res_max[0]; sr[8] = Q.res_max[1]; sr[9] = Q.res_max[2]; sr[10] = Q.res_max[3];
                sr[11] = Q.gap; sr[12] = Q.obj;
            }
            // loop condition per QP; the warp leaves when none of its QPs continues
            const bool go = kk < A.o.iter_max && Q.alpha > A.o.alpha_min
                            && (Q.res_max[0] > A.o.res_g_max || Q.res_max[1] > A.o.res_b_max || Q.res_max[2] > A.o.res_d_max
                                || Q.res_m_tau > A.o.res_m_max || Q.gap > A.o.dual_gap_max);
            if (act && !go)
            {
                int status;
                if (kk == A.o.iter_max) status = CUIPM_MAX_ITER;
                else if (Q.alpha <= A.o.alpha_min) status = CUIPM_MIN_STEP;
                else if (Q.mu != Q.mu) status = CUIPM_NAN_SOL;
                else status = CUIPM_SUCCESS;
                if (li == 0)
                {
                    info->status = status;
                    info->iter = kk;
                    for (int i = 0; i < 4; i++) info->res_max[i] = Q.res_max[i];
                    info->mu = Q.mu;
                    info->obj = Q.obj;
                    info->dual_gap = Q.gap;
                    info->lq_count = 0;
                    info->reserved = 0;
                }
                act = false;
            }
            if (!fk_any(act)) break;
            double *stt = (stat && kk + 1 < A.o.stat_max) ? stat + SM * (size_t) (kk + 1) : nullptr;
            double nrm[4] = {0, 0, 0, 0};
            double alpha = 1.0, mu_aff = 0.0, sigma_mu = 0.0;
            bool need = true;
            // affine direction: res_m already holds lam*t - tau_min (written by the residual sweep)
            for (int ph = 0; ph < 3; ph++)
            {
                const bool stw = ph < 2 ? true : need;
                if (ph == 0) fact_backward();
                else solve_backward(ph, sigma_mu, stw);
                const int do_lin = ph == 0 ? A.o.lq_fact == 1 : A.o.itref_corr_max > 0;
                double nr[4] = {0, 0, 0, 0};
                const double al = forward_pass(ph == 0, do_lin, stw, nr);
                if (stw) { alpha = al; nrm[0] = nr[0]; nrm[1] = nr[1]; nrm[2] = nr[2]; nrm[3] = nr[3]; }
                if (ph == 0)
                {
                    if (A.o.lq_fact == 1)
                    {
                        // a Cholesky step that leaves a large residual in the linear system switches the solve to the LQ
                        // refactorisation (x_ocp_qp_ipm.c:2246-2346): cold path, generic kernel
                        const double g00 = (wk + A.s0.ires.g)[0];
                        if ((nrm[0] == 0.0 && g00 != g00) || nrm[0] > 1e-5 || nrm[1] > 1e-5 || nrm[2] > 1e-5 || nrm[3] > 1e-5)
                            if (act) { hand_back(q, info); act = false; }
                    }
                    if (stt && act && li == 0) { stt[13] = 0; stt[0] = alpha; stt[1] = alpha; }
                    if (A.o.pred_corr != 1) break;
                }
                else if (ph == 2 || A.o.cond_pred_corr != 1)
                    break;
                const double mu_aff0 = mu_aff;
                mu_aff = mu_aff_pass(alpha);
                if (ph == 0)
                {
                    const double tmp = mu_aff / Q.mu;
                    const double sigma = tmp * tmp * tmp;
                    sigma_mu = sigma * Q.mu;
                    sigma_mu = sigma_mu > A.o.tau_min ? sigma_mu : A.o.tau_min;
                    if (stt && act && li == 0) { stt[2] = mu_aff; stt[3] = sigma; }
                }
                else
                {
                    need = mu_aff > 2.0 * mu_aff0;
                    if (!fk_any(act && need)) break;
                }
            }
            if (A.o.pred_corr == 1)
            {
                if (A.o.itref_corr_max > 0)
                {
                    // iterative refinement is needed when the residual of the corrector system is not small
                    // (x_ocp_qp_ipm.c:2540-2620): cold path, generic kernel
                    const bool small_ = (nrm[0] < A.o.res_g_max || nrm[0] < 1e-3 * Q.res_max[0]) && (nrm[1] < A.o.res_b_max || nrm[1] < 1e-3 * Q.res_max[1])
                                        && (nrm[2] < A.o.res_d_max || nrm[2] < 1e-3 * Q.res_max[2]) && (nrm[3] < A.o.res_m_max || nrm[3] < 1e-3 * Q.res_max[3]);
                    if (!small_ && act) { hand_back(q, info); act = false; }
                    if (stt && act && li == 0) { stt[16] = nrm[0]; stt[17] = nrm[1]; stt[18] = nrm[2]; stt[19] = nrm[3]; }
                }
                if (stt && act && li == 0) { stt[4] = alpha; stt[5] = alpha; }
            }
            if (stt && act && li == 0) stt[15] = 0;
            Q.alpha = alpha;
        }
    }

    // this QP needs a cold path: give it to the generic kernel (which starts from scratch)
    FK_DEV void hand_back(int q, cuipm_info *info)
    {
        if (li == 0)
        {
            info->status = CUIPM_FAST_REDO;
            const int slot = fk_atomic_inc(A.redo_count);
            A.redo_list[slot] = q;
        }
    }
};

// doubles of the per-QP vector pool the sweeps carve out of shared memory
inline int vector_pool_doubles(int NX, int NM, int nce, int nbe, int ns2e, int nve)
{
    const int nxe = (NX + 1) & ~1, nme = (NM + 2) & ~1;
    const int v_res = 2 * nve + 3 * nxe + 4 * nce + 2 * nbe;
    const int v_fact = 2 * nce + 2 * nbe + 3 * nme + 2 * ns2e + nxe;
    const int v_slv = nve + 2 * nce + 2 * nbe + 2 * ns2e + 2 * nxe + nme + nxe;
    const int v_fwd = 2 * nve + 6 * nxe + 8 * nce + 2 * ns2e + nbe + nme;
    const int v_init = nve + nce;
    int m = v_res;
    if (v_fact > m) m = v_fact;
    if (v_slv > m) m = v_slv;
    if (v_fwd > m) m = v_fwd;
    if (v_init > m) m = v_init;
    return m + 8;
}

}  // namespace fastk
}  // namespace cuipm
#endif
