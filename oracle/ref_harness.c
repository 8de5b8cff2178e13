#define _POSIX_C_SOURCE 199309L
/*
 * ref_harness.c -- TEST INFRASTRUCTURE ONLY.
 *
 * Thin C driver (our code) that feeds QP records in the cuipm layout (include/cuipm.h) to the UNMODIFIED
 * reference, at exactly the boundary the product replaces: the `qp_solver_config` vtable that
 * ocp_qp_hpipm_config_initialize_default() fills (acados/ocp_qp/ocp_qp_hpipm.c:517-540), i.e.
 * ocp_qp_hpipm() -> d_ocp_qp_ipm_solve().  It is compiled against the reference headers where they lie
 * (/root/reference) and linked to oracle/_ref/libacados_ref.so; see oracle/Makefile.
 *
 * Used (a) to pin oracle_ipm.c, (b) as the parity checker of the CUDA path (tests -m gpu), and (c) as the
 * CPU baseline ("kind": "reference") timed by bench.py: one solver object per OpenMP thread over a batch,
 * the structure of the reference's own batch path (c_templates_tera/acados_solver.in.c:3223-3243).
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#include "acados/ocp_qp/ocp_qp_common.h"
#include "acados/ocp_qp/ocp_qp_hpipm.h"
#include "acados/ocp_qp/ocp_qp_xcond_solver.h"
#include "acados_c/ocp_qp_interface.h"
#include "blasfeo/include/blasfeo_d_aux.h"
#include "hpipm/include/hpipm_d_ocp_qp.h"
#include "hpipm/include/hpipm_d_ocp_qp_dim.h"
#include "hpipm/include/hpipm_d_ocp_qp_ipm.h"
#include "hpipm/include/hpipm_d_ocp_qp_seed.h"
#include "hpipm/include/hpipm_d_ocp_qp_sol.h"

#include "../include/cuipm.h"
#include "oracle_ipm.h"

static double now_s(void)
{
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec + 1e-9 * ts.tv_nsec;
}

typedef struct
{
    qp_solver_config *config;
    ocp_qp_dims *dims;
    ocp_qp_in *in;
    ocp_qp_out *out;
    void *opts_mem, *mem_mem, *work_mem;
    void *opts, *mem;
} ref_obj;

static ocp_qp_dims *make_dims(const cuipm_shape *sh)
{
    ocp_qp_dims *dims = ocp_qp_dims_create(sh->N);
    for (int k = 0; k <= sh->N; k++)
    {
        int nbu = 0, nbx = 0;
        for (int i = 0; i < sh->nb[k]; i++)
            if (sh->idxb[k][i] < sh->nu[k]) nbu++; else nbx++;
        d_ocp_qp_dim_set_nx(k, sh->nx[k], dims);
        d_ocp_qp_dim_set_nu(k, sh->nu[k], dims);
        d_ocp_qp_dim_set_nbx(k, nbx, dims);
        d_ocp_qp_dim_set_nbu(k, nbu, dims);
        d_ocp_qp_dim_set_ng(k, sh->ng[k], dims);
        d_ocp_qp_dim_set_ns(k, sh->ns[k], dims);
    }
    return dims;
}

static void set_opts(ref_obj *o, const cuipm_opts *co)
{
    static const char *modes[] = {"SPEED_ABS", "SPEED", "BALANCE", "ROBUST"};
    o->config->opts_initialize_default(o->config, o->dims, o->opts);
    o->config->opts_set(o->config, o->opts, "hpipm_mode", (void *) modes[co->mode]);
    struct d_ocp_qp_ipm_arg *a = ((ocp_qp_hpipm_opts *) o->opts)->hpipm_opts;
    a->iter_max = co->iter_max; a->stat_max = co->stat_max; a->mu0 = co->mu0; a->alpha_min = co->alpha_min;
    a->res_g_max = co->res_g_max; a->res_b_max = co->res_b_max; a->res_d_max = co->res_d_max;
    a->res_m_max = co->res_m_max; a->dual_gap_max = co->dual_gap_max; a->reg_prim = co->reg_prim;
    a->lam_min = co->lam_min; a->t_min = co->t_min; a->tau_min = co->tau_min; a->lam0_min = co->lam0_min;
    a->t0_min = co->t0_min; a->pred_corr = co->pred_corr; a->cond_pred_corr = co->cond_pred_corr;
    a->itref_pred_max = co->itref_pred_max; a->itref_corr_max = co->itref_corr_max; a->lq_fact = co->lq_fact;
    a->warm_start = co->warm_start; a->abs_form = co->abs_form; a->comp_dual_sol_eq = co->comp_dual_sol_eq;
    a->comp_res_exit = co->comp_res_exit; a->split_step = co->split_step; a->var_init_scheme = co->var_init_scheme;
    a->t_lam_min = co->t_lam_min; a->t0_init = co->t0_init;
    ((ocp_qp_hpipm_opts *) o->opts)->m_relax = co->m_relax;
    o->config->opts_update(o->config, o->dims, o->opts);
}

static ref_obj *obj_create(const cuipm_shape *sh, const cuipm_opts *co)
{
    ref_obj *o = (ref_obj *) calloc(1, sizeof(ref_obj));
    o->config = (qp_solver_config *) calloc(1, sizeof(qp_solver_config));
    ocp_qp_hpipm_config_initialize_default(o->config);
    o->dims = make_dims(sh);
    o->in = ocp_qp_in_create(o->dims);
    o->out = ocp_qp_out_create(o->dims);
    o->opts_mem = calloc(1, o->config->opts_calculate_size(o->config, o->dims));
    o->opts = o->config->opts_assign(o->config, o->dims, o->opts_mem);
    set_opts(o, co);
    o->mem_mem = calloc(1, o->config->memory_calculate_size(o->config, o->dims, o->opts));
    o->mem = o->config->memory_assign(o->config, o->dims, o->opts, o->mem_mem);
    o->work_mem = calloc(1, o->config->workspace_calculate_size(o->config, o->dims, o->opts) + 8);
    for (int k = 0; k <= sh->N; k++)
    {
        if (sh->nb[k] > 0) d_ocp_qp_set_idxb(k, (int *) sh->idxb[k], o->in);
        if (sh->nb[k] + sh->ng[k] > 0 && sh->ns[k] > 0) d_ocp_qp_set_idxs_rev(k, (int *) sh->idxs_rev[k], o->in);
    }
    return o;
}

static void obj_free(ref_obj *o)
{
    o->config->terminate(o->config, o->mem, o->work_mem);
    free(o->work_mem); free(o->mem_mem); free(o->opts_mem);
    ocp_qp_out_free(o->out); ocp_qp_in_free(o->in); ocp_qp_dims_free(o->dims);
    free(o->config); free(o);
}

static void load_qp(ref_obj *o, const cuipm_shape *sh, const cuipm_layout *l, const double *q)
{
    ocp_qp_in *in = o->in;
    for (int k = 0; k <= sh->N; k++)
    {
        int n = sh->nu[k] + sh->nx[k], nc = 2 * (sh->nb[k] + sh->ng[k] + sh->ns[k]);
        if (k < sh->N)
        {
            int nx1 = sh->nx[k + 1];
            blasfeo_pack_dmat(n, nx1, (double *) q + l->off_BAt[k], n, in->BAbt + k, 0, 0);
            d_ocp_qp_set_b(k, (double *) q + l->off_b[k], in);
        }
        blasfeo_pack_dmat(n, n, (double *) q + l->off_RSQ[k], n, in->RSQrq + k, 0, 0);
        d_ocp_qp_set_r(k, (double *) q + l->off_rq[k], in);
        d_ocp_qp_set_q(k, (double *) q + l->off_rq[k] + sh->nu[k], in);
        if (sh->ng[k] > 0) blasfeo_pack_dmat(n, sh->ng[k], (double *) q + l->off_DCt[k], n, in->DCt + k, 0, 0);
        blasfeo_pack_dvec(nc, (double *) q + l->off_d[k], 1, in->d + k, 0);
        blasfeo_pack_dvec(nc, (double *) q + l->off_dmask[k], 1, in->d_mask + k, 0);
        blasfeo_dvecse(nc, 0.0, in->m + k, 0);
        if (sh->ns[k] > 0)
        {
            blasfeo_pack_dvec(2 * sh->ns[k], (double *) q + l->off_Z[k], 1, in->Z + k, 0);
            blasfeo_pack_dvec(2 * sh->ns[k], (double *) q + l->off_z[k], 1, in->rqz + k, n);
        }
    }
}

static void load_sol(ref_obj *o, const cuipm_shape *sh, const cuipm_layout *l, const double *s)
{
    for (int k = 0; k <= sh->N; k++)
    {
        int n = sh->nu[k] + sh->nx[k], nc = 2 * (sh->nb[k] + sh->ng[k] + sh->ns[k]);
        blasfeo_pack_dvec(n + 2 * sh->ns[k], (double *) s + l->off_ux[k], 1, o->out->ux + k, 0);
        if (k < sh->N) blasfeo_pack_dvec(sh->nx[k + 1], (double *) s + l->off_pi[k], 1, o->out->pi + k, 0);
        blasfeo_pack_dvec(nc, (double *) s + l->off_lam[k], 1, o->out->lam + k, 0);
        blasfeo_pack_dvec(nc, (double *) s + l->off_t[k], 1, o->out->t + k, 0);
    }
}

static void store_sol(ref_obj *o, const cuipm_shape *sh, const cuipm_layout *l, double *s)
{
    for (int k = 0; k <= sh->N; k++)
    {
        int n = sh->nu[k] + sh->nx[k], nc = 2 * (sh->nb[k] + sh->ng[k] + sh->ns[k]);
        blasfeo_unpack_dvec(n + 2 * sh->ns[k], o->out->ux + k, 0, s + l->off_ux[k], 1);
        if (k < sh->N) blasfeo_unpack_dvec(sh->nx[k + 1], o->out->pi + k, 0, s + l->off_pi[k], 1);
        blasfeo_unpack_dvec(nc, o->out->lam + k, 0, s + l->off_lam[k], 1);
        blasfeo_unpack_dvec(nc, o->out->t + k, 0, s + l->off_t[k], 1);
    }
}

/* Solve nbatch QP records with the reference.  nrep>1 repeats every solve (timing); solve_seconds returns
 * max over threads of the time spent inside config->evaluate (inputs already in the reference's own format). */
int ref_solve(const cuipm_shape *sh, int nbatch, const double *qp, double *sol, cuipm_info *info, double *stat,
              const cuipm_opts *co, int nthreads, int nrep, double *solve_seconds, double *wall_seconds)
{
    cuipm_layout *l = oracle_layout_create(sh);
    double tmax = 0.0;
#ifdef _OPENMP
    if (nthreads <= 0) nthreads = omp_get_max_threads();
#else
    nthreads = 1;
#endif
    if (nrep < 1) nrep = 1;
    double w0 = now_s();
#pragma omp parallel num_threads(nthreads)
    {
        ref_obj *o = obj_create(sh, co);
        double tacc = 0.0;
#pragma omp for schedule(dynamic, 4)
        for (int q = 0; q < nbatch; q++)
        {
            load_qp(o, sh, l, qp + (size_t) q * l->qp_stride);
            int acados_status = 0;
            for (int r = 0; r < nrep; r++)
            {
                if (co->warm_start >= 1) load_sol(o, sh, l, sol + (size_t) q * l->sol_stride);
                double t0 = now_s();
                acados_status = o->config->evaluate(o->config, o->in, o->out, o->opts, o->mem, o->work_mem);
                tacc += now_s() - t0;
            }
            (void) acados_status;
            store_sol(o, sh, l, sol + (size_t) q * l->sol_stride);
            if (info)
            {
                struct d_ocp_qp_ipm_ws *ws = ((ocp_qp_hpipm_memory *) o->mem)->hpipm_workspace;
                cuipm_info *fi = info + q;
                int st, it;
                o->config->memory_get(o->config, o->mem, "status", &st);
                o->config->memory_get(o->config, o->mem, "iter", &it);
                fi->status = st; fi->iter = it;
                d_ocp_qp_ipm_get_max_res_stat(ws, &fi->res_max[0]);
                d_ocp_qp_ipm_get_max_res_eq(ws, &fi->res_max[1]);
                d_ocp_qp_ipm_get_max_res_ineq(ws, &fi->res_max[2]);
                d_ocp_qp_ipm_get_max_res_comp(ws, &fi->res_max[3]);
                d_ocp_qp_ipm_get_obj(ws, &fi->obj);
                fi->mu = ws->res->res_mu; fi->dual_gap = ws->res->dual_gap;
                double *rstat; int sm;
                o->config->memory_get(o->config, o->mem, "stat", &rstat);
                o->config->memory_get(o->config, o->mem, "stat_m", &sm);
                int lq = 0;
                for (int i = 1; i <= it && i < co->stat_max; i++) lq += rstat[sm * i + 13] != 0.0;
                fi->lq_count = lq; fi->reserved = 0;
                if (stat)
                {
                    double *dst = stat + (size_t) q * CUIPM_STAT_M * (co->stat_max + 1);
                    memset(dst, 0, sizeof(double) * CUIPM_STAT_M * (co->stat_max + 1));
                    for (int i = 0; i <= it && i < co->stat_max; i++)
                        memcpy(dst + CUIPM_STAT_M * i, rstat + sm * i, sizeof(double) * (sm < CUIPM_STAT_M ? sm : CUIPM_STAT_M));
                }
            }
        }
#pragma omp critical
        if (tacc > tmax) tmax = tacc;
        obj_free(o);
    }
    if (wall_seconds) *wall_seconds = now_s() - w0;
    if (solve_seconds) *solve_seconds = tmax;
    oracle_layout_destroy(l);
    return 0;
}

/* Solve, then evaluate the reference's solution sensitivities (config->eval_forw_sens / eval_adj_sens, i.e.
 * d_ocp_qp_ipm_sens_frw / _adj) for one seed per QP.  seed / sens: records in the solution layout, (seed_g, seed_b,
 * seed_d, seed_m) in the (ux, pi, lam, t) slots. */
int ref_solve_sens(const cuipm_shape *sh, int nbatch, const double *qp, double *sol, cuipm_info *info, const cuipm_opts *co,
                   int nthreads, const double *seed, double *sens, int adjoint)
{
    cuipm_layout *l = oracle_layout_create(sh);
#ifdef _OPENMP
    if (nthreads <= 0) nthreads = omp_get_max_threads();
#else
    nthreads = 1;
#endif
#pragma omp parallel num_threads(nthreads)
    {
        ref_obj *o = obj_create(sh, co);
        ocp_qp_out *sout = ocp_qp_out_create(o->dims);
        void *seed_mem = calloc(1, ocp_qp_seed_calculate_size(o->dims) + 64);
        ocp_qp_seed *sd = ocp_qp_seed_assign(o->dims, seed_mem);
#pragma omp for schedule(dynamic, 4)
        for (int q = 0; q < nbatch; q++)
        {
            load_qp(o, sh, l, qp + (size_t) q * l->qp_stride);
            if (co->warm_start >= 1) load_sol(o, sh, l, sol + (size_t) q * l->sol_stride);
            o->config->evaluate(o->config, o->in, o->out, o->opts, o->mem, o->work_mem);
            store_sol(o, sh, l, sol + (size_t) q * l->sol_stride);
            if (info)
            {
                int st, it;
                o->config->memory_get(o->config, o->mem, "status", &st);
                o->config->memory_get(o->config, o->mem, "iter", &it);
                memset(info + q, 0, sizeof(cuipm_info));
                info[q].status = st; info[q].iter = it;
            }
            const double *sq = seed + (size_t) q * l->sol_stride;
            for (int k = 0; k <= sh->N; k++)
            {
                int n = sh->nu[k] + sh->nx[k], nc = 2 * (sh->nb[k] + sh->ng[k] + sh->ns[k]);
                blasfeo_pack_dvec(n + 2 * sh->ns[k], (double *) sq + l->off_ux[k], 1, sd->seed_g + k, 0);
                if (k < sh->N) blasfeo_pack_dvec(sh->nx[k + 1], (double *) sq + l->off_pi[k], 1, sd->seed_b + k, 0);
                blasfeo_pack_dvec(nc, (double *) sq + l->off_lam[k], 1, sd->seed_d + k, 0);
                blasfeo_pack_dvec(nc, (double *) sq + l->off_t[k], 1, sd->seed_m + k, 0);
            }
            if (adjoint) o->config->eval_adj_sens(o->config, o->in, sd, sout, o->opts, o->mem, o->work_mem);
            else o->config->eval_forw_sens(o->config, o->in, sd, sout, o->opts, o->mem, o->work_mem);
            ocp_qp_out *keep = o->out;
            o->out = sout;
            store_sol(o, sh, l, sens + (size_t) q * l->sol_stride);
            o->out = keep;
        }
        free(seed_mem);
        ocp_qp_out_free(sout);
        obj_free(o);
    }
    oracle_layout_destroy(l);
    return 0;
}

/* The reference's complete QP path (ocp_qp_solve -> ocp_qp_xcond_solve: stage-0 equality elimination, partial condensing to N2
 * stages, HPIPM, expansion; acados/ocp_qp/ocp_qp_xcond_solver.c:529-590) on QP records of the FULL shape: stage 0 carries x_0,
 * idxe0[0..nbxe0) are the positions (in stage 0's bound list) of the state bounds that are equalities.  sol: records of the full
 * shape.  Single-threaded (one solver object), used to pin the host-side condensing of acados_b200/ocp_qp.py. */
int ref_solve_xcond(const cuipm_shape *sh, int nbxe0, const int *idxe0, int N2, int nbatch, const double *qp, double *sol,
                    cuipm_info *info, const cuipm_opts *co)
{
    cuipm_layout *l = oracle_layout_create(sh);
    ocp_qp_solver_plan_t plan;
    plan.qp_solver = PARTIAL_CONDENSING_HPIPM;
    ocp_qp_xcond_solver_config *config = ocp_qp_xcond_solver_config_create(plan);
    ocp_qp_xcond_solver_dims *dims = ocp_qp_xcond_solver_dims_create(config, sh->N);
    for (int k = 0; k <= sh->N; k++)
    {
        int nbu = 0, nbx = 0;
        for (int i = 0; i < sh->nb[k]; i++)
            if (sh->idxb[k][i] < sh->nu[k]) nbu++; else nbx++;
        int nx = sh->nx[k], nu = sh->nu[k], ng = sh->ng[k], ns = sh->ns[k];
        config->dims_set(config, dims, k, "nx", &nx);
        config->dims_set(config, dims, k, "nu", &nu);
        config->dims_set(config, dims, k, "nbx", &nbx);
        config->dims_set(config, dims, k, "nbu", &nbu);
        config->dims_set(config, dims, k, "ng", &ng);
        config->dims_set(config, dims, k, "ns", &ns);
    }
    config->dims_set(config, dims, 0, "nbxe", &nbxe0);
    void *opts = ocp_qp_xcond_solver_opts_create(config, dims);
    ocp_qp_xcond_solver_opts_set(config, opts, "cond_N", &N2);
    int iter_max = co->iter_max, ws = co->warm_start;
    double tg = co->res_g_max, tb = co->res_b_max, td = co->res_d_max, tm = co->res_m_max;
    ocp_qp_xcond_solver_opts_set(config, opts, "iter_max", &iter_max);
    ocp_qp_xcond_solver_opts_set(config, opts, "tol_stat", &tg);
    ocp_qp_xcond_solver_opts_set(config, opts, "tol_eq", &tb);
    ocp_qp_xcond_solver_opts_set(config, opts, "tol_ineq", &td);
    ocp_qp_xcond_solver_opts_set(config, opts, "tol_comp", &tm);
    ocp_qp_xcond_solver_opts_set(config, opts, "warm_start", &ws);
    ocp_qp_solver *solver = ocp_qp_create(config, dims, opts);
    ref_obj o;
    memset(&o, 0, sizeof(o));
    o.in = ocp_qp_in_create(dims->orig_dims);
    o.out = ocp_qp_out_create(dims->orig_dims);
    int *idxbxe = (int *) calloc(nbxe0 + 1, sizeof(int));
    int nbu0 = 0;
    for (int i = 0; i < sh->nb[0]; i++) nbu0 += sh->idxb[0][i] < sh->nu[0];
    for (int e = 0; e < nbxe0; e++) idxbxe[e] = idxe0[e] - nbu0;      /* HPIPM counts within the state bounds */
    for (int k = 0; k <= sh->N; k++)
    {
        if (sh->nb[k] > 0) d_ocp_qp_set_idxb(k, (int *) sh->idxb[k], o.in);
        if (sh->nb[k] + sh->ng[k] > 0 && sh->ns[k] > 0) d_ocp_qp_set_idxs_rev(k, (int *) sh->idxs_rev[k], o.in);
    }
    if (nbxe0 > 0) d_ocp_qp_set_idxbxe(0, idxbxe, o.in);
    for (int q = 0; q < nbatch; q++)
    {
        load_qp(&o, sh, l, qp + (size_t) q * l->qp_stride);
        int st = ocp_qp_solve(solver, o.in, o.out);
        store_sol(&o, sh, l, sol + (size_t) q * l->sol_stride);
        if (info)
        {
            int it = 0;
            config->qp_solver->memory_get(config->qp_solver, ((ocp_qp_xcond_solver_memory *) solver->mem)->solver_memory, "iter", &it);
            memset(info + q, 0, sizeof(cuipm_info));
            info[q].status = st;       /* acados return value (0 success, 2 max iter, 3 min step, ...) */
            info[q].iter = it;
        }
    }
    config->terminate(config, solver->mem, solver->work);
    free(idxbxe);
    ocp_qp_out_free(o.out); ocp_qp_in_free(o.in);
    ocp_qp_solver_destroy(solver);
    ocp_qp_xcond_solver_opts_free(opts);
    ocp_qp_xcond_solver_dims_free(dims);
    ocp_qp_xcond_solver_config_free(config);
    oracle_layout_destroy(l);
    return 0;
}

int ref_num_procs(void)
{
#ifdef _OPENMP
    return omp_get_num_procs();
#else
    return 1;
#endif
}
