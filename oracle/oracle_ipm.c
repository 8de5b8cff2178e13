/*
 * oracle_ipm.c -- TEST INFRASTRUCTURE ONLY (see oracle_ipm.h).
 *
 * CPU restatement, in plain column-major C, of the reference's OCP-QP interior-point path:
 *   driver            external/hpipm/ocp_qp/x_ocp_qp_ipm.c:2684-3120   (OCP_QP_IPM_SOLVE)
 *   one iteration     external/hpipm/ocp_qp/x_ocp_qp_ipm.c:2208-2682   (OCP_QP_IPM_DELTA_STEP)
 *   starting point    external/hpipm/ocp_qp/x_ocp_qp_ipm.c:1611-2030   (OCP_QP_INIT_VAR, scheme 1)
 *   Riccati fact+solve external/hpipm/ocp_qp/x_ocp_qp_kkt.c:836-1199  (OCP_QP_FACT_SOLVE_KKT_STEP, sqrt alg.)
 *   Riccati solve     external/hpipm/ocp_qp/x_ocp_qp_kkt.c:1543-1895   (OCP_QP_SOLVE_KKT_STEP)
 *   slack elimination external/hpipm/ocp_qp/x_ocp_qp_kkt.c:220-598
 *   residuals         external/hpipm/ocp_qp/x_ocp_qp_res.c:345-727
 *   vector kernels    external/hpipm/ipm_core/x_core_qp_ipm_aux.c:38-781
 *   Cholesky pivot rule external/blasfeo/blasfeo_ref/x_lapack_ref.c:84-91 (non-positive pivot -> inverse 0)
 * with the option values acados installs (acados/ocp_qp/ocp_qp_hpipm.c:101-129).
 *
 * Parity status: PINNED against the compiled reference (oracle/_ref/libacados_ref.so built by
 * oracle/Makefile from /root/reference) by tests/test_oracle_vs_reference.py and against the committed
 * golden fixtures in tests/golden/.
 *
 * Not restated (the product rejects these option values as well): abs_form=1, split_step=1,
 * var_init_scheme=0.  info.lq_count counts the LQ refactorisations
 * (OCP_QP_FACT_LQ_SOLVE_KKT_STEP) of a solve; the Hessian factor the reference caches per solve (use_hess_fact) is
 * recomputed at each of them.
 */
#include "oracle_ipm.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* ------------------------------------------------------------------------------------------------ */
/* layout                                                                                           */
/* ------------------------------------------------------------------------------------------------ */

static size_t ev(size_t n) { return (n + 1) & ~(size_t) 1; }

cuipm_layout *oracle_layout_create(const cuipm_shape *sh)
{
    int N = sh->N;
    cuipm_layout *l = (cuipm_layout *) calloc(1, sizeof(cuipm_layout));
    size_t **arrs[] = {&l->qp_stage, &l->off_BAt, &l->off_RSQ, &l->off_DCt, &l->off_b, &l->off_rq, &l->off_d,
                       &l->off_dmask, &l->off_Z, &l->off_z, &l->sol_stage, &l->off_ux, &l->off_pi, &l->off_lam,
                       &l->off_t};
    for (unsigned i = 0; i < sizeof(arrs) / sizeof(arrs[0]); i++)
        *arrs[i] = (size_t *) calloc((size_t) N + 2, sizeof(size_t));
    l->N = N;
    size_t o = 0, s = 0;
    for (int k = 0; k <= N; k++)
    {
        size_t n = (size_t) sh->nu[k] + sh->nx[k];
        size_t nx1 = k < N ? (size_t) sh->nx[k + 1] : 0;
        size_t nc = 2 * ((size_t) sh->nb[k] + sh->ng[k] + sh->ns[k]);
        size_t ns2 = 2 * (size_t) sh->ns[k];
        l->qp_stage[k] = o;
        l->off_BAt[k] = o;   o += ev(n * nx1);
        l->off_RSQ[k] = o;   o += ev(n * n);
        l->off_DCt[k] = o;   o += ev(n * (size_t) sh->ng[k]);
        l->off_b[k] = o;     o += ev(nx1);
        l->off_rq[k] = o;    o += ev(n);
        l->off_d[k] = o;     o += ev(nc);
        l->off_dmask[k] = o; o += ev(nc);
        l->off_Z[k] = o;     o += ev(ns2);
        l->off_z[k] = o;     o += ev(ns2);
        l->sol_stage[k] = s;
        l->off_ux[k] = s;    s += ev(n + ns2);
        l->off_pi[k] = s;    s += ev(nx1);
        l->off_lam[k] = s;   s += ev(nc);
        l->off_t[k] = s;     s += ev(nc);
    }
    l->qp_stage[N + 1] = o;
    l->sol_stage[N + 1] = s;
    l->qp_stride = o;
    l->sol_stride = s;
    return l;
}

void oracle_layout_destroy(cuipm_layout *l)
{
    if (!l) return;
    free(l->qp_stage); free(l->off_BAt); free(l->off_RSQ); free(l->off_DCt); free(l->off_b); free(l->off_rq);
    free(l->off_d); free(l->off_dmask); free(l->off_Z); free(l->off_z); free(l->sol_stage); free(l->off_ux);
    free(l->off_pi); free(l->off_lam); free(l->off_t);
    free(l);
}

/* ------------------------------------------------------------------------------------------------ */
/* per-QP work area                                                                                 */
/* ------------------------------------------------------------------------------------------------ */

typedef struct { double **ux, **pi, **lam, **t; } vset;   /* a primal-dual point or step */
typedef struct { double **g, **b, **d, **m; } rset;       /* a residual / right-hand side */

typedef struct
{
    int N;
    const int *nx, *nu, *nb, *ng, *ns;
    const int *const *idxb;
    const int *const *idxs_rev;
    int *nv, *nc;                 /* nu+nx, 2nb+2ng+2ns */
    int nct;                      /* total number of constraints */
    /* data (pointers into the QP record) */
    const double **BAt, **RSQ, **DCt, **b, **rq, **d, **dmask, **Z, **z;
    vset sol, step, itref;
    rset res, res_itref;
    double **res_m_bkp;
    double **lam_bkp, **t_bkp;     /* iterate of the last factorisation (UPDATE_VAR backups, x_core_qp_ipm_aux.c:534-575) */
    double **L, **Linv, **lrow, **Pb, **Gamma, **gamma, **t_inv, **Zs_inv;
    double *AL;                   /* (nvmax+1) x (nxmax + ngmax) scratch */
    double *lq;                   /* nvmax x (2 nvmax + ngmax + nxmax) scratch of the LQ refactorisation */
    double *tmp0, *tmp1, *tmp2, *tmp3, *tmpx, *tmpl; /* nb+ng / nx scratch */
    int mask_constr;
    double m_relax;     /* entries of qp->m (all equal) */
    double m_safe;      /* share of m the ratio test keeps lam*t above (x_ocp_qp_ipm.c:104-209: 0.3 SPEED modes, 0.5 BALANCE / ROBUST) */
    double nc_mask_inv;
    char *arena;
} work;

static double *bump(char **p, size_t n)
{
    double *r = (double *) *p;
    *p += ((n + 1) & ~(size_t) 1) * sizeof(double);
    return r;
}

static double **bumpp(char **p, size_t n)
{
    double **r = (double **) *p;
    *p += n * sizeof(double *);
    return r;
}

static work *work_create(const cuipm_shape *sh)
{
    int N = sh->N;
    size_t nvt = 0, net = 0, nct = 0, nLt = 0;
    int nvmax = 0, nxmax = 0, ngmax = 0, nbgmax = 0;
    for (int k = 0; k <= N; k++)
    {
        int n = sh->nu[k] + sh->nx[k];
        int nc = 2 * (sh->nb[k] + sh->ng[k] + sh->ns[k]);
        nvt += (size_t) n + 2 * sh->ns[k] + 2;
        net += (size_t) (k < N ? sh->nx[k + 1] : 0) + 2;
        nct += (size_t) nc + 2;
        nLt += (size_t) n * n + 2 * n + 4;
        if (n > nvmax) nvmax = n;
        if (sh->nx[k] > nxmax) nxmax = sh->nx[k];
        if (sh->ng[k] > ngmax) ngmax = sh->ng[k];
        if (sh->nb[k] + sh->ng[k] > nbgmax) nbgmax = sh->nb[k] + sh->ng[k];
    }
    size_t bytes = sizeof(double) * (4 * nvt + 4 * net + 18 * nct + nLt + (size_t) (nvmax + 1) * (nxmax + ngmax + 2)
                                     + 4 * (size_t) (nbgmax + 2) + 2 * (size_t) (nvmax + 2) + 64)
                   + sizeof(double *) * 64 * (size_t) (N + 2) + sizeof(int) * 2 * (size_t) (N + 2);
    work *w = (work *) calloc(1, sizeof(work));
    w->arena = (char *) calloc(1, bytes);
    char *p = w->arena;
    w->N = N; w->nx = sh->nx; w->nu = sh->nu; w->nb = sh->nb; w->ng = sh->ng; w->ns = sh->ns;
    w->idxb = sh->idxb; w->idxs_rev = sh->idxs_rev;
    size_t np = (size_t) N + 1;
    double ***pp[] = {(double ***) &w->BAt, (double ***) &w->RSQ, (double ***) &w->DCt, (double ***) &w->b,
                      (double ***) &w->rq, (double ***) &w->d, (double ***) &w->dmask, (double ***) &w->Z,
                      (double ***) &w->z, &w->sol.ux, &w->sol.pi, &w->sol.lam, &w->sol.t, &w->step.ux, &w->step.pi,
                      &w->step.lam, &w->step.t, &w->itref.ux, &w->itref.pi, &w->itref.lam, &w->itref.t, &w->res.g,
                      &w->res.b, &w->res.d, &w->res.m, &w->res_itref.g, &w->res_itref.b, &w->res_itref.d,
                      &w->res_itref.m, &w->res_m_bkp, &w->lam_bkp, &w->t_bkp, &w->L, &w->Linv, &w->lrow, &w->Pb, &w->Gamma, &w->gamma,
                      &w->t_inv, &w->Zs_inv};
    for (unsigned i = 0; i < sizeof(pp) / sizeof(pp[0]); i++) *pp[i] = bumpp(&p, np);
    w->nv = (int *) p; p += sizeof(int) * (np + (np & 1));
    w->nc = (int *) p; p += sizeof(int) * (np + (np & 1));
    w->nct = 0;
    for (int k = 0; k <= N; k++)
    {
        int n = sh->nu[k] + sh->nx[k];
        int nc = 2 * (sh->nb[k] + sh->ng[k] + sh->ns[k]);
        int nx1 = k < N ? sh->nx[k + 1] : 0;
        int nvs = n + 2 * sh->ns[k];
        w->nv[k] = n; w->nc[k] = nc; w->nct += nc;
        w->step.ux[k] = bump(&p, nvs); w->itref.ux[k] = bump(&p, nvs);
        w->res.g[k] = bump(&p, nvs); w->res_itref.g[k] = bump(&p, nvs);
        w->step.pi[k] = bump(&p, nx1); w->itref.pi[k] = bump(&p, nx1);
        w->res.b[k] = bump(&p, nx1); w->res_itref.b[k] = bump(&p, nx1);
        w->step.lam[k] = bump(&p, nc); w->step.t[k] = bump(&p, nc);
        w->itref.lam[k] = bump(&p, nc); w->itref.t[k] = bump(&p, nc);
        w->res.d[k] = bump(&p, nc); w->res.m[k] = bump(&p, nc);
        w->res_itref.d[k] = bump(&p, nc); w->res_itref.m[k] = bump(&p, nc);
        w->res_m_bkp[k] = bump(&p, nc);
        w->lam_bkp[k] = bump(&p, nc); w->t_bkp[k] = bump(&p, nc);
        w->Gamma[k] = bump(&p, nc); w->gamma[k] = bump(&p, nc); w->t_inv[k] = bump(&p, nc);
        w->Zs_inv[k] = bump(&p, 2 * sh->ns[k]);
        w->L[k] = bump(&p, (size_t) n * n); w->Linv[k] = bump(&p, n); w->lrow[k] = bump(&p, n);
        w->Pb[k] = bump(&p, nx1);
    }
    w->AL = bump(&p, (size_t) (nvmax + 1) * (nxmax + ngmax + 1));
    w->tmp0 = bump(&p, nbgmax); w->tmp1 = bump(&p, nbgmax); w->tmp2 = bump(&p, nbgmax); w->tmp3 = bump(&p, nbgmax);
    w->tmpx = bump(&p, nvmax + 1); w->tmpl = bump(&p, nvmax + 1);
    w->lq = (double *) calloc((size_t) nvmax * (2 * nvmax + ngmax + nxmax) + 2, sizeof(double));
    return w;
}

static void work_destroy(work *w)
{
    free(w->lq);
    free(w->arena);
    free(w);
}

static void work_bind(work *w, const cuipm_layout *l, const double *qp, double *sol)
{
    for (int k = 0; k <= w->N; k++)
    {
        w->BAt[k] = qp + l->off_BAt[k]; w->RSQ[k] = qp + l->off_RSQ[k]; w->DCt[k] = qp + l->off_DCt[k];
        w->b[k] = qp + l->off_b[k]; w->rq[k] = qp + l->off_rq[k]; w->d[k] = qp + l->off_d[k];
        w->dmask[k] = qp + l->off_dmask[k]; w->Z[k] = qp + l->off_Z[k]; w->z[k] = qp + l->off_z[k];
        w->sol.ux[k] = sol + l->off_ux[k]; w->sol.pi[k] = sol + l->off_pi[k];
        w->sol.lam[k] = sol + l->off_lam[k]; w->sol.t[k] = sol + l->off_t[k];
    }
}

/* ------------------------------------------------------------------------------------------------ */
/* residuals  (x_ocp_qp_res.c:345-531 nonlinear, :535-683 linearised)                               */
/* ------------------------------------------------------------------------------------------------ */

/* y = tril-symmetric(H) x, H n x n column-major lower */
static void symv_l(int n, const double *H, const double *x, double *y)
{
    for (int i = 0; i < n; i++) y[i] = 0.0;
    for (int j = 0; j < n; j++)
    {
        y[j] += H[j + n * j] * x[j];
        for (int i = j + 1; i < n; i++)
        {
            y[i] += H[i + n * j] * x[j];
            y[j] += H[i + n * j] * x[i];
        }
    }
}

/* Common body.  lin==0: residuals of the QP at point `p` (rqz,b,d from the data; m = 0).
 * lin==1: residuals of the Newton system with right-hand side `rhs` at step `p`, linearised at `at`. */
static void res_body(work *w, int lin, const vset *p, const rset *rhs, const vset *at, rset *res, double *mu_out,
                     double *obj_out, double *gap_out)
{
    int N = w->N;
    double mu = 0.0, obj = 0.0, gap = 0.0;
    for (int k = 0; k <= N; k++)
    {
        int nu0 = w->nu[k], nx0 = w->nx[k], n = nu0 + nx0, nb0 = w->nb[k], ng0 = w->ng[k], ns0 = w->ns[k];
        int nbg = nb0 + ng0, nc = w->nc[k];
        const double *ux = p->ux[k], *lam = p->lam[k], *t = p->t[k], *mask = w->dmask[k];
        const double *gvec = lin ? rhs->g[k] : w->rq[k];
        const double *zvec = lin ? rhs->g[k] + n : w->z[k];
        const double *dvec = lin ? rhs->d[k] : w->d[k];
        double *rg = res->g[k], *rd = res->d[k], *rm = res->m[k];

        symv_l(n, w->RSQ[k], ux, rg);
        if (!lin)
        {
            double s1 = 0.0, s2 = 0.0;
            for (int i = 0; i < n; i++) { rg[i] += 2.0 * gvec[i]; s1 += rg[i] * ux[i]; }
            obj += 0.5 * s1;
            for (int i = 0; i < n; i++) { rg[i] -= gvec[i]; s2 += rg[i] * ux[i]; }
            gap += s2;
        }
        else
            for (int i = 0; i < n; i++) rg[i] += gvec[i];
        if (k > 0)
            for (int i = 0; i < nx0; i++) rg[nu0 + i] -= p->pi[k - 1][i];

        /* masked multipliers are formed on the fly: lmask(i) = lam[i]*mask[i] */
#define LMASK(i) (w->mask_constr ? lam[i] * mask[i] : lam[i])
        if (nbg > 0)
        {
            for (int i = 0; i < nbg; i++) w->tmp0[i] = LMASK(nbg + i) - LMASK(i);
            for (int i = 0; i < 2 * nbg; i++) rd[i] = t[i] + dvec[i];
            for (int i = 0; i < nb0; i++)
            {
                rg[w->idxb[k][i]] += w->tmp0[i];
                w->tmp1[i] = ux[w->idxb[k][i]];
            }
            if (ng0 > 0)
            {
                const double *C = w->DCt[k];
                for (int g = 0; g < ng0; g++)
                {
                    double acc = 0.0, lg = w->tmp0[nb0 + g];
                    for (int i = 0; i < n; i++)
                    {
                        rg[i] += C[i + n * g] * lg;
                        acc += C[i + n * g] * ux[i];
                    }
                    w->tmp1[nb0 + g] = acc;
                }
            }
            for (int i = 0; i < nbg; i++)
            {
                rd[i] -= w->tmp1[i];
                rd[nbg + i] += w->tmp1[i];
            }
        }
        if (ns0 > 0)
        {
            const double *Z = w->Z[k];
            double *rgs = rg + n;
            const double *s = ux + n;
            if (!lin)
            {
                double s1 = 0.0, s2 = 0.0;
                for (int i = 0; i < 2 * ns0; i++) { rgs[i] = Z[i] * s[i] + 2.0 * zvec[i]; s1 += rgs[i] * s[i]; }
                obj += 0.5 * s1;
                for (int i = 0; i < 2 * ns0; i++) { rgs[i] -= zvec[i]; s2 += rgs[i] * s[i]; }
                gap += s2;
            }
            else
                for (int i = 0; i < 2 * ns0; i++) rgs[i] = Z[i] * s[i] + zvec[i];
            for (int i = 0; i < 2 * ns0; i++) rgs[i] -= LMASK(2 * nbg + i);
            for (int i = 0; i < nbg; i++)
            {
                int idx = w->idxs_rev[k][i];
                if (idx != -1)
                {
                    rgs[idx] -= LMASK(i);
                    rgs[ns0 + idx] -= LMASK(nbg + i);
                    rd[i] -= s[idx];
                    rd[nbg + i] -= s[ns0 + idx];
                }
            }
            for (int i = 0; i < 2 * ns0; i++) rd[2 * nbg + i] = t[2 * nbg + i] - s[i] + dvec[2 * nbg + i];
        }
        if (w->mask_constr)
            for (int i = 0; i < nc; i++) rd[i] *= mask[i];
        if (!lin)
            for (int i = 0; i < nc; i++) gap -= dvec[i] * LMASK(i);

        if (k < N)
        {
            int nx1 = w->nx[k + 1], nu1 = w->nu[k + 1];
            const double *A = w->BAt[k], *bv = lin ? rhs->b[k] : w->b[k], *pik = p->pi[k];
            double *rb = res->b[k];
            for (int j = 0; j < nx1; j++)
            {
                double acc = 0.0, pj = pik[j];
                for (int i = 0; i < n; i++)
                {
                    rg[i] += A[i + n * j] * pj;
                    acc += A[i + n * j] * ux[i];
                }
                rb[j] = bv[j] - p->ux[k + 1][nu1 + j] + acc;
            }
            if (!lin)
                for (int j = 0; j < nx1; j++) gap -= bv[j] * pik[j];
        }

        if (!lin)
        {
            double tmu = 0.0;
            for (int i = 0; i < nc; i++)
            {
                rm[i] = LMASK(i) * t[i] - w->m_relax;       /* qp->m = m_relax everywhere (ocp_qp_hpipm.c:338-342, x_ocp_qp_res.c:513-514) */
                if (w->mask_constr) rm[i] *= mask[i];
                tmu += fabs(rm[i]);
            }
            mu += tmu;
        }
        else
        {
            const double *Lam = at->lam[k], *T = at->t[k], *mv = rhs->m[k];
            for (int i = 0; i < nc; i++)
            {
                rm[i] = mv[i] + Lam[i] * t[i] + lam[i] * T[i];
                if (w->mask_constr) rm[i] *= mask[i];
            }
        }
#undef LMASK
    }
    if (mu_out) *mu_out = mu * w->nc_mask_inv;
    if (obj_out) *obj_out = obj;
    if (gap_out) *gap_out = gap;
}

static void res_inf_norm(const work *w, const rset *res, double out[4])
{
    int nan[4] = {0, 0, 0, 0};
    out[0] = out[1] = out[2] = out[3] = 0.0;
    for (int k = 0; k <= w->N; k++)
    {
        int nvs = w->nv[k] + 2 * w->ns[k], nx1 = k < w->N ? w->nx[k + 1] : 0, nc = w->nc[k];
        /* VECNRM_INF semantics of BLASFEO (auxiliary/d_aux_lib4.c:4893-4995): max of fabs, NaN if any entry is NaN */
        for (int i = 0; i < nvs; i++) { double a = fabs(res->g[k][i]); if (a > out[0]) out[0] = a; if (a != a) nan[0] = 1; }
        for (int i = 0; i < nx1; i++) { double a = fabs(res->b[k][i]); if (a > out[1]) out[1] = a; if (a != a) nan[1] = 1; }
        for (int i = 0; i < nc; i++) { double a = fabs(res->d[k][i]); if (a > out[2]) out[2] = a; if (a != a) nan[2] = 1; }
        for (int i = 0; i < nc; i++) { double a = fabs(res->m[k][i]); if (a > out[3]) out[3] = a; if (a != a) nan[3] = 1; }
    }
    for (int i = 0; i < 4; i++)
        if (nan[i]) out[i] = NAN;
}

/* ------------------------------------------------------------------------------------------------ */
/* slack elimination (x_ocp_qp_kkt.c:220-335, 431-520, 524-598)                                     */
/* ------------------------------------------------------------------------------------------------ */

/* fact!=0: also (re)compute Zs_inv and tmp0 (effective Gamma).  Always: slack part of dux and tmp1 (effective gamma). */
static void cond_slacks(work *w, int k, int fact, const rset *rhs, vset *step, double reg)
{
    int n = w->nv[k], nb0 = w->nb[k], ng0 = w->ng[k], ns0 = w->ns[k], nbg = nb0 + ng0;
    const double *Gam = w->Gamma[k], *gam = w->gamma[k], *Z = w->Z[k];
    double *Zi = w->Zs_inv[k], *ds = step->ux[k] + n;
    const double *rgs = rhs->g[k] + n;
    const int *rev = w->idxs_rev[k];
    for (int j = 0; j < ns0; j++)
    {
        if (fact)
        {
            Zi[j] = Z[j] + reg + Gam[2 * nbg + j];
            Zi[ns0 + j] = Z[ns0 + j] + reg + Gam[2 * nbg + ns0 + j];
        }
        ds[j] = rgs[j] + gam[2 * nbg + j];
        ds[ns0 + j] = rgs[ns0 + j] + gam[2 * nbg + ns0 + j];
    }
    for (int i = 0; i < nbg; i++)
    {
        int j = rev[i];
        if (j != -1)
        {
            if (fact) { Zi[j] += Gam[i]; Zi[ns0 + j] += Gam[nbg + i]; }
            ds[j] += gam[i];
            ds[ns0 + j] += gam[nbg + i];
        }
    }
    if (fact)
        for (int j = 0; j < 2 * ns0; j++) Zi[j] = 1.0 / Zi[j];
    for (int i = 0; i < nbg; i++)
    {
        int j = rev[i];
        if (j != -1)
        {
            if (fact)
            {
                w->tmp0[i] = Gam[i] - Gam[i] * Zi[j] * Gam[i];
                w->tmp1[i] = Gam[nbg + i] - Gam[nbg + i] * Zi[ns0 + j] * Gam[nbg + i];
            }
            w->tmp2[i] = gam[i] - Gam[i] * Zi[j] * ds[j];
            w->tmp3[i] = gam[nbg + i] - Gam[nbg + i] * Zi[ns0 + j] * ds[ns0 + j];
        }
        else
        {
            if (fact) { w->tmp0[i] = Gam[i]; w->tmp1[i] = Gam[nbg + i]; }
            w->tmp2[i] = gam[i];
            w->tmp3[i] = gam[nbg + i];
        }
    }
    for (int i = 0; i < nbg; i++)
    {
        if (fact) w->tmp0[i] = w->tmp0[i] + w->tmp1[i];
        w->tmp1[i] = w->tmp2[i] - w->tmp3[i];
    }
}

static void expand_slacks(work *w, int k, vset *step)
{
    int n = w->nv[k], nb0 = w->nb[k], ng0 = w->ng[k], ns0 = w->ns[k], nbg = nb0 + ng0;
    const double *Gam = w->Gamma[k], *Zi = w->Zs_inv[k];
    double *ds = step->ux[k] + n, *dt = step->t[k];
    const int *rev = w->idxs_rev[k];
    for (int i = 0; i < nbg; i++)
    {
        int j = rev[i];
        if (j != -1)
        {
            ds[j] += Gam[i] * dt[i];
            ds[ns0 + j] += Gam[nbg + i] * dt[nbg + i];
        }
    }
    for (int j = 0; j < ns0; j++)
    {
        ds[j] = -Zi[j] * ds[j];
        ds[ns0 + j] = -Zi[ns0 + j] * ds[ns0 + j];
        dt[2 * nbg + j] = ds[j];
        dt[2 * nbg + ns0 + j] = ds[ns0 + j];
    }
    for (int i = 0; i < nbg; i++)
    {
        int j = rev[i];
        if (j != -1)
        {
            dt[i] += ds[j];
            dt[nbg + i] += ds[ns0 + j];
        }
    }
}

/* ------------------------------------------------------------------------------------------------ */
/* Riccati                                                                                          */
/* ------------------------------------------------------------------------------------------------ */

/* In-place lower Cholesky of the n x n column-major M, with the row vector `row` (length n) carried along:
 * on exit M = L, Linv[j] = 1/L[j][j] (0 for a non-positive pivot), row = row * L^{-T}.
 * Pivot rule: blasfeo_ref/x_lapack_ref.c:84-91. */
static void potrf_row(int n, double *M, double *Linv, double *row)
{
    for (int j = 0; j < n; j++)
    {
        double c = M[j + n * j];
        for (int kk = 0; kk < j; kk++) c -= M[j + n * kk] * M[j + n * kk];
        double inv = c > 0.0 ? 1.0 / sqrt(c) : 0.0;
        Linv[j] = inv;
        M[j + n * j] = c * inv;
        for (int i = j + 1; i < n; i++)
        {
            double a = M[i + n * j];
            for (int kk = 0; kk < j; kk++) a -= M[i + n * kk] * M[j + n * kk];
            M[i + n * j] = a * inv;
        }
        double r = row[j];
        for (int kk = 0; kk < j; kk++) r -= row[kk] * M[j + n * kk];
        row[j] = r * inv;
    }
}

/* after the forward/backward sweeps: dt from dux, slack expansion, dlam/dt (x_ocp_qp_kkt.c:1176-1193) */
static void finish_step(work *w, const rset *rhs, vset *step, int mask_out)
{
    for (int k = 0; k <= w->N; k++)
    {
        int n = w->nv[k], nb0 = w->nb[k], ng0 = w->ng[k], nbg = nb0 + ng0, nc = w->nc[k];
        double *dt = step->t[k], *dlam = step->lam[k];
        const double *dux = step->ux[k];
        for (int i = 0; i < nb0; i++) dt[i] = dux[w->idxb[k][i]];
        for (int g = 0; g < ng0; g++)
        {
            double acc = 0.0;
            for (int i = 0; i < n; i++) acc += w->DCt[k][i + n * g] * dux[i];
            dt[nb0 + g] = acc;
        }
        for (int i = 0; i < nbg; i++) dt[nbg + i] = -dt[i];
        if (w->ns[k] > 0) expand_slacks(w, k, step);
        /* COMPUTE_LAM_T_QP, x_core_qp_ipm_aux.c:164-189 */
        const double *lam = w->sol.lam[k], *ti = w->t_inv[k], *rd = rhs->d[k], *rm = rhs->m[k];
        for (int i = 0; i < nc; i++)
        {
            dlam[i] = -ti[i] * (rm[i] + (lam[i] * dt[i]) - (lam[i] * rd[i]));
            dt[i] -= rd[i];
        }
        /* the reference masks the step of the main direction only: after an iterative-refinement solve it
         * re-masks sol_step instead of sol_itref (x_ocp_qp_ipm.c:2599-2606), so the correction is added unmasked */
        if (w->mask_constr && mask_out)
            for (int i = 0; i < nc; i++) { dt[i] *= w->dmask[k][i]; dlam[i] *= w->dmask[k][i]; }
    }
}

/* forward substitution shared by fact_solve and solve: on entry step->ux[k][0:nu] (stage 0: [0:nv]) holds the
 * negated backward quantity, step->pi[k] holds the "p" part (or 0); x_{k+1} is produced on the fly. */
static void forward_sweep(work *w, const rset *rhs, vset *step, int pi_has_p)
{
    int N = w->N;
    for (int k = 0; k <= N; k++)
    {
        int nu0 = w->nu[k], n = w->nv[k];
        const double *L = w->L[k], *Li = w->Linv[k];
        double *v = step->ux[k];
        int nsolve = k == 0 ? n : nu0;
        /* TRSV_LTN(_MN): v[0:nsolve] = L[0:n,0:nsolve]^{-T} (v[0:nsolve] - L[nsolve:n,0:nsolve]' v[nsolve:n]) */
        for (int j = nsolve - 1; j >= 0; j--)
        {
            double a = v[j];
            for (int i = j + 1; i < n; i++) a -= L[i + n * j] * v[i];
            v[j] = a * Li[j];
        }
        if (k < N)
        {
            int nx1 = w->nx[k + 1], nu1 = w->nu[k + 1], n1 = w->nv[k + 1];
            const double *A = w->BAt[k], *L1 = w->L[k + 1];
            double *x1 = step->ux[k + 1] + nu1, *pi = step->pi[k];
            for (int j = 0; j < nx1; j++)
            {
                double acc = rhs->b[k][j];
                for (int i = 0; i < n; i++) acc += A[i + n * j] * v[i];
                x1[j] = acc;
            }
            /* pi = p + Lxx (Lxx' x1),  Lxx = L1[nu1:, nu1:] */
            for (int j = 0; j < nx1; j++)
            {
                double acc = 0.0;
                for (int i = j; i < nx1; i++) acc += L1[(nu1 + i) + n1 * (nu1 + j)] * x1[i];
                w->tmpx[j] = acc;
            }
            if (!pi_has_p)
                for (int j = 0; j < nx1; j++) w->tmpx[j] += w->lrow[k + 1][nu1 + j];
            for (int i = nx1 - 1; i >= 0; i--)
            {
                double acc = 0.0;
                for (int j = 0; j <= i; j++) acc += L1[(nu1 + i) + n1 * (nu1 + j)] * w->tmpx[j];
                pi[i] = pi_has_p ? pi[i] + acc : acc;
            }
        }
    }
}

static void compute_Gamma_gamma(work *w, const rset *rhs, int with_Gamma, int t_lam_min, double lam_min, double t_min)
{
    double t_min_inv = t_min > 0 ? 1.0 / t_min : 1e30;
    for (int k = 0; k <= w->N; k++)
    {
        const double *lam = w->sol.lam[k], *t = w->sol.t[k], *rd = rhs->d[k], *rm = rhs->m[k];
        for (int i = 0; i < w->nc[k]; i++)
        {
            if (with_Gamma)
            {
                w->t_inv[k][i] = 1.0 / t[i];
                if (t_lam_min == 1)
                {
                    double ti = t[i] < t_min ? t_min_inv : w->t_inv[k][i];
                    double lt = lam[i] < lam_min ? lam_min : lam[i];
                    w->Gamma[k][i] = ti * lt;
                }
                else
                    w->Gamma[k][i] = w->t_inv[k][i] * lam[i];
            }
            w->gamma[k][i] = w->t_inv[k][i] * (rm[i] - lam[i] * rd[i]);
        }
    }
}

/* OCP_QP_FACT_SOLVE_KKT_STEP, square-root algorithm (x_ocp_qp_kkt.c:880-1007) */
static void fact_solve_kkt_step(work *w, const rset *rhs, vset *step, const cuipm_opts *o)
{
    int N = w->N;
    compute_Gamma_gamma(w, rhs, 1, o->t_lam_min, o->lam_min, o->t_min);
    for (int k = N; k >= 0; k--)
    {
        int nu0 = w->nu[k], n = w->nv[k], nb0 = w->nb[k], ng0 = w->ng[k], nbg = nb0 + ng0;
        (void) nu0;
        double *M = w->L[k], *row = w->lrow[k];
        const double *H = w->RSQ[k];
        for (int j = 0; j < n; j++)
        {
            for (int i = 0; i < j; i++) M[i + n * j] = 0.0;
            for (int i = j; i < n; i++) M[i + n * j] = H[i + n * j];
            M[j + n * j] += o->reg_prim;
            row[j] = rhs->g[k][j];
        }
        if (w->ns[k] > 0)
            cond_slacks(w, k, 1, rhs, step, o->reg_prim);
        else
            for (int i = 0; i < nbg; i++)
            {
                w->tmp0[i] = w->Gamma[k][i] + w->Gamma[k][nbg + i];
                w->tmp1[i] = w->gamma[k][i] - w->gamma[k][nbg + i];
            }
        for (int i = 0; i < nb0; i++)
        {
            int ix = w->idxb[k][i];
            M[ix + n * ix] += w->tmp0[i];
            row[ix] += w->tmp1[i];
        }
        int kc = 0;                 /* columns of AL accumulated into the syrk */
        int ldal = n + 1;
        double *AL = w->AL;
        if (k < N)
        {
            int nx1 = w->nx[k + 1], nu1 = w->nu[k + 1], n1 = w->nv[k + 1];
            const double *A = w->BAt[k], *L1 = w->L[k + 1];
            /* AL = [A; res_b'] * Lxx  (TRMM_RLNN) */
            for (int j = 0; j < nx1; j++)
            {
                for (int i = 0; i < n; i++)
                {
                    double acc = 0.0;
                    for (int c = j; c < nx1; c++) acc += A[i + n * c] * L1[(nu1 + c) + n1 * (nu1 + j)];
                    AL[i + ldal * j] = acc;
                }
                double acc = 0.0;
                for (int c = j; c < nx1; c++) acc += rhs->b[k][c] * L1[(nu1 + c) + n1 * (nu1 + j)];
                AL[n + ldal * j] = acc;
            }
            /* Pb = Lxx * (Lxx' b) */
            for (int i = nx1 - 1; i >= 0; i--)
            {
                double acc = 0.0;
                for (int j = 0; j <= i; j++) acc += L1[(nu1 + i) + n1 * (nu1 + j)] * AL[n + ldal * j];
                w->Pb[k][i] = acc;
            }
            for (int j = 0; j < nx1; j++) AL[n + ldal * j] += w->lrow[k + 1][nu1 + j];
            kc = nx1;
        }
        /* M += AL AL' (lower), row += ALrow AL' */
        for (int c = 0; c < kc; c++)
            for (int j = 0; j < n; j++)
            {
                double ajc = AL[j + ldal * c];
                for (int i = j; i < n; i++) M[i + n * j] += AL[i + ldal * c] * ajc;
                row[j] += AL[n + ldal * c] * ajc;
            }
        /* general constraints: M += C diag(tmp0) C', row += tmp1' C' */
        for (int g = 0; g < ng0; g++)
        {
            const double *C = w->DCt[k] + (size_t) n * g;
            double Gg = w->tmp0[nb0 + g], gg = w->tmp1[nb0 + g];
            for (int j = 0; j < n; j++)
            {
                double cj = C[j];
                for (int i = j; i < n; i++) M[i + n * j] += C[i] * Gg * cj;
                row[j] += gg * cj;
            }
        }
        potrf_row(n, M, w->Linv[k], row);
    }
    /* forward */
    for (int k = 0; k <= N; k++)
    {
        int nneg = k == 0 ? w->nv[k] : w->nu[k];
        for (int i = 0; i < nneg; i++) step->ux[k][i] = -w->lrow[k][i];
    }
    forward_sweep(w, rhs, step, 0);
    finish_step(w, rhs, step, 1);
}

/* Householder LQ with non-negative diagonal of the n x m column-major matrix X (ld n), in place: on exit the lower
 * triangle of the first n columns holds L with X X' = L L'.  Reflector formulas of BLASFEO's positive-diagonal
 * kernels (kernel/generic/kernel_dgeqrf_4_lib4.c:4743-4771: beta = +sqrt(sigma + alpha^2), the pivot is left alone
 * when the rest of the row is zero), unblocked. */
static void gelqf_pd(int n, int m, double *X)
{
    for (int i = 0; i < n; i++)
    {
        double sigma = 0.0;
        for (int j = i + 1; j < m; j++) sigma += X[i + n * j] * X[i + n * j];
        if (sigma == 0.0) continue;
        double alpha = X[i + n * i];
        double beta = sqrt(sigma + alpha * alpha);
        double tmp = alpha <= 0 ? alpha - beta : -sigma / (alpha + beta);
        double tau = 2 * tmp * tmp / (sigma + tmp * tmp);
        tmp = 1.0 / tmp;
        X[i + n * i] = beta;
        for (int j = i + 1; j < m; j++) X[i + n * j] *= tmp;
        for (int r = i + 1; r < n; r++)
        {
            double ww = X[r + n * i];
            for (int j = i + 1; j < m; j++) ww += X[r + n * j] * X[i + n * j];
            ww = -ww * tau;
            X[r + n * i] += ww;
            for (int j = i + 1; j < m; j++) X[r + n * j] += ww * X[i + n * j];
        }
    }
}

/* OCP_QP_FACT_LQ_SOLVE_KKT_STEP (x_ocp_qp_kkt.c:1201-1541): L_k from an LQ factorisation of
 * [chol(RSQ_k + reg I) | sqrt(Gamma_b) on idxb | DCt sqrt(Gamma_g) | BAt L_{k+1,xx}] instead of the Cholesky
 * factorisation of the accumulated Gram matrix; the gradient goes through plain substitutions. */
static void fact_lq_solve_kkt_step(work *w, const rset *rhs, vset *step, const cuipm_opts *o)
{
    int N = w->N;
    compute_Gamma_gamma(w, rhs, 1, o->t_lam_min, o->lam_min, o->t_min);
    for (int k = N; k >= 0; k--)
    {
        int nu0 = w->nu[k], n = w->nv[k], nb0 = w->nb[k], ng0 = w->ng[k], nbg = nb0 + ng0;
        int nx1 = k < N ? w->nx[k + 1] : 0;
        int m = 2 * n + ng0 + nx1;
        double *X = w->lq, *v = step->ux[k];
        memset(X, 0, sizeof(double) * (size_t) n * m);
        for (int i = 0; i < n; i++) v[i] = rhs->g[k][i];
        if (k < N)
        {
            int nu1 = w->nu[k + 1], n1 = w->nv[k + 1];
            const double *A = w->BAt[k], *L1 = w->L[k + 1];
            double *AL = X + (size_t) n * (2 * n + ng0);
            for (int j = 0; j < nx1; j++)
                for (int i = 0; i < n; i++)
                {
                    double acc = 0.0;
                    for (int c = j; c < nx1; c++) acc += A[i + n * c] * L1[(nu1 + c) + n1 * (nu1 + j)];
                    AL[i + n * j] = acc;
                }
            for (int j = 0; j < nx1; j++)
            {
                double acc = 0.0;
                for (int i = j; i < nx1; i++) acc += L1[(nu1 + i) + n1 * (nu1 + j)] * rhs->b[k][i];
                w->tmpl[j] = acc;
            }
            for (int i = nx1 - 1; i >= 0; i--)
            {
                double acc = 0.0;
                for (int j = 0; j <= i; j++) acc += L1[(nu1 + i) + n1 * (nu1 + j)] * w->tmpl[j];
                w->Pb[k][i] = acc;
            }
            for (int j = 0; j < nx1; j++) w->tmpx[j] = step->ux[k + 1][nu1 + j] + w->Pb[k][j];
            for (int j = 0; j < nx1; j++)
                for (int i = 0; i < n; i++) v[i] += A[i + n * j] * w->tmpx[j];
        }
        if (w->ns[k] > 0)
            cond_slacks(w, k, 1, rhs, step, o->reg_prim);
        else
            for (int i = 0; i < nbg; i++)
            {
                w->tmp0[i] = w->Gamma[k][i] + w->Gamma[k][nbg + i];
                w->tmp1[i] = w->gamma[k][i] - w->gamma[k][nbg + i];
            }
        for (int i = 0; i < nb0; i++)
        {
            int ix = w->idxb[k][i];
            double t = w->tmp0[i] >= 0.0 ? w->tmp0[i] : 0.0;
            t = sqrt(t);
            X[ix + n * (n + ix)] = t > 0.0 ? t : 0.0;
            v[ix] += w->tmp1[i];
        }
        for (int g = 0; g < ng0; g++)
        {
            const double *C = w->DCt[k] + (size_t) n * g;
            double t = w->tmp0[nb0 + g] >= 0.0 ? w->tmp0[nb0 + g] : 0.0;
            t = sqrt(t);
            for (int i = 0; i < n; i++)
            {
                X[i + n * (2 * n + g)] = C[i] * t;
                v[i] += C[i] * w->tmp1[nb0 + g];
            }
        }
        /* Lh = chol(tril(RSQ) + reg I) (the reference caches it per solve: use_hess_fact) */
        {
            const double *H = w->RSQ[k];
            for (int j = 0; j < n; j++)
            {
                for (int i = j; i < n; i++) X[i + n * j] = H[i + n * j];
                X[j + n * j] += o->reg_prim;
            }
            for (int j = 0; j < n; j++) w->tmpl[j] = 0.0;
            potrf_row(n, X, w->Linv[k], w->tmpl);
        }
        gelqf_pd(n, m, X);
        double *L = w->L[k], *Li = w->Linv[k];
        for (int j = 0; j < n; j++)
        {
            for (int i = 0; i < j; i++) L[i + n * j] = 0.0;
            for (int i = j; i < n; i++) L[i + n * j] = X[i + n * j];
            Li[j] = 1.0 / L[j + n * j];
        }
        int nsolve = k == 0 ? n : nu0;
        for (int j = 0; j < nsolve; j++)
        {
            double a = v[j];
            for (int c = 0; c < j; c++) a -= L[j + n * c] * v[c];
            v[j] = a * Li[j];
        }
        for (int i = nsolve; i < n; i++)
        {
            double a = v[i];
            for (int c = 0; c < nsolve; c++) a -= L[i + n * c] * v[c];
            v[i] = a;
        }
    }
    for (int k = 0; k <= N; k++)
    {
        if (k < N)
            for (int j = 0; j < w->nx[k + 1]; j++) step->pi[k][j] = step->ux[k + 1][w->nu[k + 1] + j];
        int nneg = k == 0 ? w->nv[k] : w->nu[k];
        for (int i = 0; i < nneg; i++) step->ux[k][i] = -step->ux[k][i];
    }
    forward_sweep(w, rhs, step, 1);
    finish_step(w, rhs, step, 1);
}

/* OCP_QP_SOLVE_KKT_STEP, square-root algorithm (x_ocp_qp_kkt.c:1582-1727) */
static void solve_kkt_step(work *w, const rset *rhs, vset *step, int use_Pb, int mask_out)
{
    int N = w->N;
    compute_Gamma_gamma(w, rhs, 0, 0, 0, 0);
    for (int k = N; k >= 0; k--)
    {
        int nu0 = w->nu[k], n = w->nv[k], nb0 = w->nb[k], ng0 = w->ng[k], nbg = nb0 + ng0;
        double *v = step->ux[k];
        for (int i = 0; i < n; i++) v[i] = rhs->g[k][i];
        if (w->ns[k] > 0)
            cond_slacks(w, k, 0, rhs, step, 0.0);
        else
            for (int i = 0; i < nbg; i++) w->tmp1[i] = w->gamma[k][i] - w->gamma[k][nbg + i];
        for (int i = 0; i < nb0; i++) v[w->idxb[k][i]] += w->tmp1[i];
        for (int g = 0; g < ng0; g++)
            for (int i = 0; i < n; i++) v[i] += w->DCt[k][i + n * g] * w->tmp1[nb0 + g];
        if (k < N)
        {
            int nx1 = w->nx[k + 1], nu1 = w->nu[k + 1], n1 = w->nv[k + 1];
            const double *A = w->BAt[k], *L1 = w->L[k + 1];
            if (use_Pb)
                for (int j = 0; j < nx1; j++) w->tmpx[j] = step->ux[k + 1][nu1 + j] + w->Pb[k][j];
            else
            {
                for (int j = 0; j < nx1; j++)
                {
                    double acc = 0.0;
                    for (int i = j; i < nx1; i++) acc += L1[(nu1 + i) + n1 * (nu1 + j)] * rhs->b[k][i];
                    w->tmpl[j] = acc;
                }
                for (int i = nx1 - 1; i >= 0; i--)
                {
                    double acc = 0.0;
                    for (int j = 0; j <= i; j++) acc += L1[(nu1 + i) + n1 * (nu1 + j)] * w->tmpl[j];
                    w->tmpx[i] = acc + step->ux[k + 1][nu1 + i];
                }
            }
            for (int j = 0; j < nx1; j++)
                for (int i = 0; i < n; i++) v[i] += A[i + n * j] * w->tmpx[j];
        }
        /* TRSV_LNN(_MN) */
        int nsolve = k == 0 ? n : nu0;
        const double *L = w->L[k], *Li = w->Linv[k];
        for (int j = 0; j < nsolve; j++)
        {
            double a = v[j];
            for (int c = 0; c < j; c++) a -= L[j + n * c] * v[c];
            v[j] = a * Li[j];
        }
        for (int i = nsolve; i < n; i++)
        {
            double a = v[i];
            for (int c = 0; c < nsolve; c++) a -= L[i + n * c] * v[c];
            v[i] = a;
        }
    }
    /* forward: pi_k starts as the backward value of x_{k+1}; the unknown part of v is negated */
    for (int k = 0; k <= N; k++)
    {
        if (k < N)
            for (int j = 0; j < w->nx[k + 1]; j++) step->pi[k][j] = step->ux[k + 1][w->nu[k + 1] + j];
        int nneg = k == 0 ? w->nv[k] : w->nu[k];
        for (int i = 0; i < nneg; i++) step->ux[k][i] = -step->ux[k][i];
    }
    forward_sweep(w, rhs, step, 1);
    finish_step(w, rhs, step, mask_out);
}

/* ------------------------------------------------------------------------------------------------ */
/* IPM vector kernels (x_core_qp_ipm_aux.c)                                                         */
/* ------------------------------------------------------------------------------------------------ */

/* COMPUTE_ALPHA_QP, single step: m == 0 (:375-398) and m != 0 (:398-440: the step is also limited by lam*t >= m_safe*m, the
 * smaller root of the quadratic; every test uses the step length the tests before it left) */
static double compute_alpha(const work *w, const vset *step)
{
    double alpha = 1.0;
    for (int k = 0; k <= w->N; k++)
    {
        const double *lam = w->sol.lam[k], *t = w->sol.t[k], *dlam = step->lam[k], *dt = step->t[k], *mask = w->dmask[k];
        for (int i = 0; i < w->nc[k]; i++)
        {
            if (w->m_relax == 0.0)
            {
                if (lam[i] + alpha * dlam[i] < 0.0) alpha = -lam[i] / dlam[i];
                if (t[i] + alpha * dt[i] < 0.0) alpha = -t[i] / dt[i];
                continue;
            }
            double lam1 = lam[i] + alpha * dlam[i], t1 = t[i] + alpha * dt[i];
            if (lam1 < 0.0) { alpha = -lam[i] / dlam[i]; lam1 = lam[i] + alpha * dlam[i]; }
            if (t1 < 0.0) { alpha = -t[i] / dt[i]; t1 = t[i] + alpha * dt[i]; }
            const double m1 = w->m_safe * w->m_relax * mask[i];         /* cws->m = qp->m * d_mask (x_ocp_qp_ipm.c:2722) */
            if (lam1 * t1 - m1 < -1e-12)
            {
                const double c = lam[i] * t[i] - m1;
                if (c > 0.0)
                {
                    const double a = dlam[i] * dt[i], b = dlam[i] * t[i] + lam[i] * dt[i];
                    const double d = b * b - 4.0 * a * c, sd = sqrt(d), tmp = 0.5 / a;
                    alpha = (-b - sd) * tmp;
                }
                else
                    alpha = 0.0;
            }
        }
    }
    return alpha;
}

/* COMPUTE_MU_AFF_QP (:636-668) */
static double compute_mu_aff(const work *w, const vset *step, double alpha)
{
    double mu = 0.0;
    for (int k = 0; k <= w->N; k++)
    {
        const double *lam = w->sol.lam[k], *t = w->sol.t[k], *dlam = step->lam[k], *dt = step->t[k];
        /* (:636-668: |(lam + alpha dlam)(t + alpha dt) - m|, m = qp->m * d_mask) */
        if (w->m_relax == 0.0)
            for (int i = 0; i < w->nc[k]; i++) mu += fabs((lam[i] + alpha * dlam[i]) * (t[i] + alpha * dt[i]));
        else
            for (int i = 0; i < w->nc[k]; i++) mu += fabs(-w->m_relax * w->dmask[k][i] + (lam[i] + alpha * dlam[i]) * (t[i] + alpha * dt[i]));
    }
    return mu * w->nc_mask_inv;
}

static void mask_res_m(work *w)
{
    if (!w->mask_constr) return;
    for (int k = 0; k <= w->N; k++)
        for (int i = 0; i < w->nc[k]; i++) w->res.m[k][i] *= w->dmask[k][i];
}

/* UPDATE_VAR_QP (:472-582) */
static void update_var(work *w, double alpha, const cuipm_opts *o)
{
    if (alpha < 1.0) alpha = alpha * ((1.0 - alpha) * 0.99 + alpha * 0.9999999);
    for (int k = 0; k <= w->N; k++)
    {
        int nvs = w->nv[k] + 2 * w->ns[k], nx1 = k < w->N ? w->nx[k + 1] : 0, nc = w->nc[k];
        for (int i = 0; i < nvs; i++) w->sol.ux[k][i] += alpha * w->step.ux[k][i];
        for (int i = 0; i < nx1; i++) w->sol.pi[k][i] += alpha * w->step.pi[k][i];
        for (int i = 0; i < nc; i++)
        {
            w->lam_bkp[k][i] = w->sol.lam[k][i];
            w->t_bkp[k][i] = w->sol.t[k][i];
            double l = w->sol.lam[k][i] + alpha * w->step.lam[k][i];
            double t = w->sol.t[k][i] + alpha * w->step.t[k][i];
            if (o->t_lam_min == 2)
            {
                l = l <= o->lam_min ? o->lam_min : l;
                t = t <= o->t_min ? o->t_min : t;
            }
            w->sol.lam[k][i] = l;
            w->sol.t[k][i] = t;
        }
        if (w->mask_constr)
            for (int i = 0; i < nc; i++) w->sol.lam[k][i] *= w->dmask[k][i];
    }
}

/* OCP_QP_INIT_VAR (x_ocp_qp_ipm.c:1611-2030), var_init_scheme 1 */
static void init_var(work *w, const cuipm_opts *o)
{
    const double thr0 = 0.1;
    int N = w->N;
    /* the plugin zeroes the primal iterate before every solve, whatever warm_start says
     * (acados/ocp_qp/ocp_qp_hpipm.c:333-336): warm starts carry over pi, lam and t only */
    for (int k = 0; k <= N; k++)
        for (int i = 0; i < w->nv[k] + 2 * w->ns[k]; i++) w->sol.ux[k][i] = 0.0;
    if (o->warm_start >= 2)
    {
        double lmin = o->warm_start >= 3 ? o->lam0_min : thr0, tmin = o->warm_start >= 3 ? o->t0_min : thr0;
        for (int k = 0; k <= N; k++)
            for (int i = 0; i < w->nc[k]; i++)
            {
                if (w->sol.lam[k][i] < lmin) w->sol.lam[k][i] = lmin;
                if (w->sol.t[k][i] < tmin) w->sol.t[k][i] = tmin;
            }
        return;
    }
    for (int k = 0; k < N; k++)
        for (int i = 0; i < w->nx[k + 1]; i++) w->sol.pi[k][i] = 0.0;
    if (o->t0_init == 0 || o->t0_init == 1)
    {
        double l0 = o->t0_init == 0 ? sqrt(o->mu0) : o->mu0, t0 = o->t0_init == 0 ? sqrt(o->mu0) : 1.0;
        for (int k = 0; k <= N; k++)
            for (int i = 0; i < w->nc[k]; i++) { w->sol.lam[k][i] = l0; w->sol.t[k][i] = t0; }
        return;
    }
    for (int k = 0; k <= N; k++)
    {
        int n = w->nv[k], nb0 = w->nb[k], ng0 = w->ng[k], ns0 = w->ns[k], nbg = nb0 + ng0;
        double *ux = w->sol.ux[k], *s = ux + n, *t = w->sol.t[k], *lam = w->sol.lam[k];
        const double *d = w->d[k];
        const int *idxb = w->idxb[k], *rev = w->idxs_rev ? w->idxs_rev[k] : 0;
        for (int j = 0; j < 2 * ns0; j++)
        {
            double tj = s[j] - d[2 * nbg + j];
            if (tj < thr0) { tj = thr0; s[j] = d[2 * nbg + j] + tj; }
            t[2 * nbg + j] = tj;
        }
        for (int j = 0; j < nb0; j++)
        {
            double tl = ux[idxb[j]], tu = -ux[idxb[j]];
            if (ns0 > 0 && rev[j] != -1) { tl += s[rev[j]]; tu += s[ns0 + rev[j]]; }
            tl -= d[j];
            tu -= d[nbg + j];
            if (tl < thr0)
            {
                if (tu < thr0)
                {
                    ux[idxb[j]] = 0.5 * (d[j] - d[nbg + j]);
                    tl = thr0; tu = thr0;
                }
                else
                {
                    tl = thr0;
                    ux[idxb[j]] = d[j] + thr0;
                }
            }
            else if (tu < thr0)
            {
                tu = thr0;
                ux[idxb[j]] = -d[nbg + j] - thr0;
            }
            t[j] = tl; t[nbg + j] = tu;
        }
        for (int g = 0; g < ng0; g++)
        {
            double acc = 0.0;
            for (int i = 0; i < n; i++) acc += w->DCt[k][i + n * g] * ux[i];
            double tl = acc, tu = -acc;
            if (ns0 > 0 && rev[nb0 + g] != -1) { tl += s[rev[nb0 + g]]; tu += s[ns0 + rev[nb0 + g]]; }
            tl -= d[nb0 + g];
            tu -= d[nbg + nb0 + g];
            t[nb0 + g] = thr0 > tl ? thr0 : tl;
            t[nbg + nb0 + g] = thr0 > tu ? thr0 : tu;
        }
        for (int i = 0; i < w->nc[k]; i++) lam[i] = o->mu0 / t[i];
    }
}

/* ------------------------------------------------------------------------------------------------ */
/* driver                                                                                           */
/* ------------------------------------------------------------------------------------------------ */

static void add_step(work *w, vset *dst, const vset *src)
{
    for (int k = 0; k <= w->N; k++)
    {
        int nvs = w->nv[k] + 2 * w->ns[k], nx1 = k < w->N ? w->nx[k + 1] : 0, nc = w->nc[k];
        for (int i = 0; i < nvs; i++) dst->ux[k][i] += src->ux[k][i];
        for (int i = 0; i < nx1; i++) dst->pi[k][i] += src->pi[k][i];
        for (int i = 0; i < nc; i++) { dst->lam[k][i] += src->lam[k][i]; dst->t[k][i] += src->t[k][i]; }
    }
}

static int itref_ok(const double nrm[4], const double resmax[4], const cuipm_opts *o)
{
    return (nrm[0] < o->res_g_max || nrm[0] < 1e-3 * resmax[0]) && (nrm[1] < o->res_b_max || nrm[1] < 1e-3 * resmax[1])
           && (nrm[2] < o->res_d_max || nrm[2] < 1e-3 * resmax[2]) && (nrm[3] < o->res_m_max || nrm[3] < 1e-3 * resmax[3]);
}

static void solve_one(work *w, const cuipm_opts *o, cuipm_info *info, double *stat)
{
    w->m_relax = o->m_relax;
    w->m_safe = (o->mode == CUIPM_SPEED_ABS || o->mode == CUIPM_SPEED) ? 0.3 : 0.5;
    int N = w->N;
    const int SM = CUIPM_STAT_M;
    double res_max[4] = {0, 0, 0, 0}, mu = 0, obj = 0, gap = 0;
    int lq_count = 0, force_lq = 0;
    if (stat) memset(stat, 0, sizeof(double) * SM * (size_t) (o->stat_max + 1));

    /* constraint masks (x_ocp_qp_ipm.c:2774-2806) */
    int nc_mask = 0;
    for (int k = 0; k <= N; k++)
        for (int i = 0; i < w->nc[k]; i++)
            if (w->dmask[k][i] != 0.0) nc_mask++;
    w->mask_constr = nc_mask < w->nct;
    w->nc_mask_inv = nc_mask > 0 ? 1.0 / nc_mask : 0.0;

    if (w->nct == 0 || nc_mask == 0)
    {
        /* unconstrained: one Riccati pass on the QP data itself (OCP_QP_FACT_SOLVE_KKT_UNCONSTR, x_ocp_qp_kkt.c:39-216) */
        rset data = {(double **) w->rq, (double **) w->b, w->res.d, w->res.m};
        /* rq has no slack part in this case only if ns==0; slack-only problems keep z in res.g */
        for (int k = 0; k <= N; k++)
        {
            for (int i = 0; i < w->nc[k]; i++) { w->res.d[k][i] = 0.0; w->res.m[k][i] = 0.0; w->sol.lam[k][i] = 0.0; w->sol.t[k][i] = 1.0; }
            for (int i = 0; i < w->nv[k]; i++) w->res.g[k][i] = w->rq[k][i];
            for (int i = 0; i < 2 * w->ns[k]; i++) w->res.g[k][w->nv[k] + i] = w->z[k][i];
        }
        data.g = w->res.g;
        fact_solve_kkt_step(w, &data, &w->step, o);
        for (int k = 0; k <= N; k++)
        {
            int nvs = w->nv[k] + 2 * w->ns[k], nx1 = k < N ? w->nx[k + 1] : 0;
            for (int i = 0; i < nvs; i++) w->sol.ux[k][i] = w->step.ux[k][i];
            for (int i = 0; i < nx1; i++) w->sol.pi[k][i] = w->step.pi[k][i];
            for (int i = 0; i < w->nc[k]; i++) { w->sol.lam[k][i] = 0.0; }
        }
        res_body(w, 0, &w->sol, 0, 0, &w->res, &mu, &obj, &gap);
        res_inf_norm(w, &w->res, res_max);
        if (stat && 0 < o->stat_max)
        {   /* column quirk of the reference's unconstrained branch (x_ocp_qp_ipm.c:2822-2829) */
            stat[6] = res_max[0]; stat[7] = res_max[1]; stat[8] = res_max[2]; stat[9] = res_max[3];
            stat[10] = gap; stat[11] = obj;
        }
        info->status = isnan(w->sol.ux[0][0]) ? CUIPM_NAN_SOL : CUIPM_SUCCESS;
        info->iter = 0;
        goto fill;
    }

    init_var(w, o);
    if (w->mask_constr)
        for (int k = 0; k <= N; k++)
            for (int i = 0; i < w->nc[k]; i++) w->sol.lam[k][i] *= w->dmask[k][i];

    double alpha = 1.0;
    res_body(w, 0, &w->sol, 0, 0, &w->res, &mu, &obj, &gap);
    res_inf_norm(w, &w->res, res_max);
    if (stat && 0 < o->stat_max)
    {
        stat[7] = res_max[0]; stat[8] = res_max[1]; stat[9] = res_max[2]; stat[10] = res_max[3];
        stat[11] = gap; stat[12] = obj;
    }
#define RES_M_TAU(out)                                                                                   \
    do {                                                                                                 \
        double r_ = 0.0;                                                                                 \
        for (int k = 0; k <= N; k++)                                                                     \
            for (int i = 0; i < w->nc[k]; i++)                                                           \
            {                                                                                            \
                double a_ = fabs(w->res.m[k][i] - o->tau_min * w->dmask[k][i]);                          \
                if (a_ > r_ || a_ != a_) r_ = a_;                                                        \
            }                                                                                            \
        out = r_;                                                                                        \
    } while (0)
    double res_m_tau;
    RES_M_TAU(res_m_tau);

    int kk;
    for (kk = 0; kk < o->iter_max && alpha > o->alpha_min
                 && (res_max[0] > o->res_g_max || res_max[1] > o->res_b_max || res_max[2] > o->res_d_max
                     || res_m_tau > o->res_m_max || gap > o->dual_gap_max);
         kk++)
    {
        double *st = (stat && kk + 1 < o->stat_max) ? stat + SM * (size_t) (kk + 1) : 0;
        double nrm[4] = {0, 0, 0, 0};
        /* affine direction: res_m <- res_m - tau_min (x_ocp_qp_ipm.c:2236-2244) */
        for (int k = 0; k <= N; k++)
            for (int i = 0; i < w->nc[k]; i++)
            {
                w->res_m_bkp[k][i] = w->res.m[k][i];
                w->res.m[k][i] = w->res_m_bkp[k][i] - o->tau_min;
            }
        mask_res_m(w);
        /* factorisation: Cholesky, with a switch to LQ for the rest of the solve once the linear-system residual of
         * a Cholesky step is too large (x_ocp_qp_ipm.c:2246-2346) */
        if (o->lq_fact == 0 || (o->lq_fact == 1 && !force_lq))
        {
            fact_solve_kkt_step(w, &w->res, &w->step, o);
            if (st) st[13] = 0;
            if (o->lq_fact == 1)
            {
                res_body(w, 1, &w->step, &w->res, &w->sol, &w->res_itref, 0, 0, 0);
                res_inf_norm(w, &w->res_itref, nrm);
                if ((nrm[0] == 0.0 && isnan(w->res_itref.g[0][0])) || nrm[0] > 1e-5 || nrm[1] > 1e-5 || nrm[2] > 1e-5
                    || nrm[3] > 1e-5)
                {
                    fact_lq_solve_kkt_step(w, &w->res, &w->step, o);
                    force_lq = 1;
                    lq_count++;
                    if (st) st[13] = 1;
                }
            }
        }
        else
        {
            fact_lq_solve_kkt_step(w, &w->res, &w->step, o);
            lq_count++;
            if (st) st[13] = 1;
        }
        alpha = compute_alpha(w, &w->step);
        if (st) { st[0] = alpha; st[1] = alpha; }
        int itref1 = 0;
        if (o->pred_corr == 1)
        {
            double mu_aff = compute_mu_aff(w, &w->step, alpha);
            double tmp = mu_aff / mu;
            double sigma = tmp * tmp * tmp;
            double sigma_mu = sigma * mu;
            sigma_mu = sigma_mu > o->tau_min ? sigma_mu : o->tau_min;
            if (st) { st[2] = mu_aff; st[3] = sigma; }
            /* centering + correction (x_core_qp_ipm_aux.c:695-722) */
            for (int k = 0; k <= N; k++)
                for (int i = 0; i < w->nc[k]; i++)
                    w->res.m[k][i] = w->res_m_bkp[k][i] + w->step.t[k][i] * w->step.lam[k][i] - sigma_mu;
            mask_res_m(w);
            solve_kkt_step(w, &w->res, &w->step, 1, 1);
            alpha = compute_alpha(w, &w->step);
            if (o->cond_pred_corr == 1)
            {
                double mu_aff0 = mu_aff;
                mu_aff = compute_mu_aff(w, &w->step, alpha);
                if (mu_aff > 2.0 * mu_aff0)
                {
                    for (int k = 0; k <= N; k++)
                        for (int i = 0; i < w->nc[k]; i++) w->res.m[k][i] = w->res_m_bkp[k][i] - sigma_mu;
                    mask_res_m(w);
                    solve_kkt_step(w, &w->res, &w->step, 1, 1);
                    alpha = compute_alpha(w, &w->step);
                }
            }
            int iter_ref_step = 0;
            if (o->itref_corr_max > 0)
            {
                for (itref1 = 0; itref1 < o->itref_corr_max; itref1++)
                {
                    res_body(w, 1, &w->step, &w->res, &w->sol, &w->res_itref, 0, 0, 0);
                    res_inf_norm(w, &w->res_itref, nrm);
                    if (itref_ok(nrm, res_max, o)) break;
                    solve_kkt_step(w, &w->res_itref, &w->itref, 0, 0);
                    iter_ref_step = 1;
                    add_step(w, &w->step, &w->itref);
                }
                if (itref1 == o->itref_corr_max)
                {
                    res_body(w, 1, &w->step, &w->res, &w->sol, &w->res_itref, 0, 0, 0);
                    res_inf_norm(w, &w->res_itref, nrm);
                }
                if (st) { st[16] = nrm[0]; st[17] = nrm[1]; st[18] = nrm[2]; st[19] = nrm[3]; }
            }
            if (iter_ref_step) alpha = compute_alpha(w, &w->step);
            if (st) { st[4] = alpha; st[5] = alpha; }
        }
        if (st) st[15] = itref1;
        update_var(w, alpha, o);

        res_body(w, 0, &w->sol, 0, 0, &w->res, &mu, &obj, &gap);
        res_inf_norm(w, &w->res, res_max);
        if (st)
        {
            st[6] = mu; st[7] = res_max[0]; st[8] = res_max[1]; st[9] = res_max[2]; st[10] = res_max[3];
            st[11] = gap; st[12] = obj;
        }
        RES_M_TAU(res_m_tau);
    }
    info->iter = kk;
    if (kk == o->iter_max) info->status = CUIPM_MAX_ITER;
    else if (alpha <= o->alpha_min) info->status = CUIPM_MIN_STEP;
    else if (isnan(mu)) info->status = CUIPM_NAN_SOL;
    else info->status = CUIPM_SUCCESS;
fill:
    for (int i = 0; i < 4; i++) info->res_max[i] = res_max[i];
    info->mu = mu; info->obj = obj; info->dual_gap = gap; info->lq_count = lq_count; info->reserved = 0;
}

static int opts_supported(const cuipm_opts *o)
{
    return o->abs_form == 0 && o->split_step == 0 && o->comp_dual_sol_eq == 1 && o->comp_res_exit == 1
           && o->var_init_scheme == 1 && o->itref_pred_max == 0;
}

int oracle_solve(const cuipm_shape *shape, int nbatch, const double *qp, double *sol, cuipm_info *info, double *stat,
                 const cuipm_opts *opts, int nthreads)
{
    if (!opts_supported(opts)) return CUIPM_ERR_INVALID;
    cuipm_layout *l = oracle_layout_create(shape);
#ifdef _OPENMP
    if (nthreads <= 0) nthreads = omp_get_max_threads();
#else
    nthreads = 1;
#endif
#pragma omp parallel num_threads(nthreads)
    {
        work *w = work_create(shape);
#pragma omp for schedule(dynamic, 4)
        for (int q = 0; q < nbatch; q++)
        {
            work_bind(w, l, qp + (size_t) q * l->qp_stride, sol + (size_t) q * l->sol_stride);
            solve_one(w, opts, info + q, stat ? stat + (size_t) q * CUIPM_STAT_M * (opts->stat_max + 1) : 0);
        }
        work_destroy(w);
    }
    oracle_layout_destroy(l);
    return CUIPM_OK;
}

/* OCP_QP_IPM_SENS_FRW / _ADJ (x_ocp_qp_ipm.c:3285-3444): one substitution with the factorisation of the last IPM
 * iteration, at the iterate that factorisation was computed at (the UPDATE_VAR backups), the seed being the right-hand
 * side.  seed / sens use the solution-record layout: (seed_g, seed_b, seed_d, seed_m) in the (ux, pi, lam, t) slots. */
static void sens_one(work *w, const cuipm_layout *l, const double *seed, double *sens, int adjoint, double *scratch)
{
    int N = w->N;
    rset rhs;
    vset out;
    double **pp = (double **) calloc(8 * (size_t) (N + 1), sizeof(double *));
    rhs.g = pp; rhs.b = pp + (N + 1); rhs.d = pp + 2 * (N + 1); rhs.m = pp + 3 * (N + 1);
    out.ux = pp + 4 * (N + 1); out.pi = pp + 5 * (N + 1); out.lam = pp + 6 * (N + 1); out.t = pp + 7 * (N + 1);
    memcpy(scratch, seed, sizeof(double) * l->sol_stride);
    for (int k = 0; k <= N; k++)
    {
        rhs.g[k] = scratch + l->off_ux[k]; rhs.b[k] = scratch + l->off_pi[k];
        rhs.d[k] = scratch + l->off_lam[k]; rhs.m[k] = scratch + l->off_t[k];
        out.ux[k] = sens + l->off_ux[k]; out.pi[k] = sens + l->off_pi[k];
        out.lam[k] = sens + l->off_lam[k]; out.t[k] = sens + l->off_t[k];
        if (adjoint)
            for (int i = 0; i < w->nc[k]; i++) rhs.m[k][i] *= w->t_bkp[k][i];
    }
    double **lam_cur = w->sol.lam, **t_cur = w->sol.t;
    w->sol.lam = w->lam_bkp; w->sol.t = w->t_bkp;
    solve_kkt_step(w, &rhs, &out, 0, 0);
    w->sol.lam = lam_cur; w->sol.t = t_cur;
    if (adjoint)
        for (int k = 0; k <= N; k++)
            for (int i = 0; i < w->nc[k]; i++) out.t[k][i] *= w->t_inv[k][i];
    free(pp);
}

int oracle_solve_sens(const cuipm_shape *shape, int nbatch, const double *qp, double *sol, cuipm_info *info,
                      const cuipm_opts *opts, int nthreads, const double *seed, double *sens, int adjoint)
{
    if (!opts_supported(opts)) return CUIPM_ERR_INVALID;
    cuipm_layout *l = oracle_layout_create(shape);
#ifdef _OPENMP
    if (nthreads <= 0) nthreads = omp_get_max_threads();
#else
    nthreads = 1;
#endif
#pragma omp parallel num_threads(nthreads)
    {
        work *w = work_create(shape);
        double *scratch = (double *) calloc(l->sol_stride + 2, sizeof(double));
#pragma omp for schedule(dynamic, 4)
        for (int q = 0; q < nbatch; q++)
        {
            work_bind(w, l, qp + (size_t) q * l->qp_stride, sol + (size_t) q * l->sol_stride);
            solve_one(w, opts, info + q, 0);
            sens_one(w, l, seed + (size_t) q * l->sol_stride, sens + (size_t) q * l->sol_stride, adjoint, scratch);
        }
        free(scratch);
        work_destroy(w);
    }
    oracle_layout_destroy(l);
    return CUIPM_OK;
}

int oracle_residuals(const cuipm_shape *shape, int nbatch, const double *qp, const double *sol, cuipm_info *info)
{
    cuipm_layout *l = oracle_layout_create(shape);
    work *w = work_create(shape);
    for (int q = 0; q < nbatch; q++)
    {
        work_bind(w, l, qp + (size_t) q * l->qp_stride, (double *) sol + (size_t) q * l->sol_stride);
        int nc_mask = 0;
        for (int k = 0; k <= w->N; k++)
            for (int i = 0; i < w->nc[k]; i++)
                if (w->dmask[k][i] != 0.0) nc_mask++;
        w->mask_constr = nc_mask < w->nct;
        w->nc_mask_inv = nc_mask > 0 ? 1.0 / nc_mask : 0.0;
        res_body(w, 0, &w->sol, 0, 0, &w->res, &info[q].mu, &info[q].obj, &info[q].dual_gap);
        res_inf_norm(w, &w->res, info[q].res_max);
    }
    work_destroy(w);
    oracle_layout_destroy(l);
    return CUIPM_OK;
}
