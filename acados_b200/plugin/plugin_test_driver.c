/*
 * plugin_test_driver.c -- TEST DRIVER (our code): runs the reference's ocp_qp_xcond_solver with its own HPIPM plugin
 * and with the cuipm plugin swapped into the same slot, on the reference's mass-spring problem class
 * (examples/c/no_interface_examples/mass_spring_model/mass_spring_qp.c: nx=8, nu=3, N=15, nb=11, x0 as stage-0
 * equality boxes eliminated by the partial condensing module), for N2 in {N, 5, 3} as test/ocp_qp/test_qpsolvers.cpp
 * does, plus the batched entry, the Riccati getters and the memory_get fields.  Exit code 0 = all checks passed.
 *
 * The swap `ocp_qp_cuipm_config_initialize_default(config->qp_solver)` is exactly what a
 * `case PARTIAL_CONDENSING_CUIPM:` in ocp_qp_xcond_solver_config_initialize_from_plan would do.
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "acados/ocp_qp/ocp_qp_common.h"
#include "acados/ocp_qp/ocp_qp_xcond_solver.h"
#include "acados_c/ocp_qp_interface.h"
#include "blasfeo/include/blasfeo_d_aux.h"
#include "hpipm/include/hpipm_d_ocp_qp.h"

#include "acados/ocp_qp/ocp_qp_hpipm.h"
#include "hpipm/include/hpipm_d_ocp_qp_dim.h"

#include "ocp_qp_cuipm.h"

#define NX 8
#define NU 3
#define NN 15

static void mass_spring(double Ts, double *A, double *B)
{   /* exp of the augmented matrix [[Ac, Bc], [0, 0]] * Ts by a Taylor series (norm ~ 1.5) */
    enum { M = NX + NU };
    double Mx[M * M], T[M * M], E[M * M], T2[M * M];
    int pp = NX / 2;
    memset(Mx, 0, sizeof(Mx));
    for (int i = 0; i < pp; i++)
    {
        Mx[i + M * (pp + i)] = Ts;                 /* p' = v */
        Mx[(pp + i) + M * i] = -2 * Ts;            /* v' = T p + u */
        if (i > 0) Mx[(pp + i) + M * (i - 1)] = Ts;
        if (i < pp - 1) Mx[(pp + i) + M * (i + 1)] = Ts;
    }
    for (int i = 0; i < NU; i++) Mx[(pp + i) + M * (NX + i)] = Ts;
    memset(E, 0, sizeof(E)); memset(T, 0, sizeof(T));
    for (int i = 0; i < M; i++) E[i + M * i] = T[i + M * i] = 1.0;
    for (int it = 1; it < 30; it++)
    {
        memset(T2, 0, sizeof(T2));
        for (int j = 0; j < M; j++) for (int k = 0; k < M; k++) for (int i = 0; i < M; i++) T2[i + M * j] += T[i + M * k] * Mx[k + M * j] / it;
        memcpy(T, T2, sizeof(T));
        for (int i = 0; i < M * M; i++) E[i] += T[i];
    }
    for (int j = 0; j < NX; j++) for (int i = 0; i < NX; i++) A[i + NX * j] = E[i + M * j];
    for (int j = 0; j < NU; j++) for (int i = 0; i < NX; i++) B[i + NX * j] = E[i + M * (NX + j)];
}

typedef struct { ocp_qp_xcond_solver_config *config; ocp_qp_xcond_solver_dims *dims; void *opts; ocp_qp_solver *solver; } chain;

static chain make_chain(int use_cuipm, int N2)
{
    chain c;
#ifdef CUIPM_REGISTERED
    /* a libacados built with ACADOS_WITH_CUIPM (integration/Makefile): the solver is selected by name, like any other */
    c.config = ocp_qp_xcond_solver_config_create_from_name(use_cuipm ? "PARTIAL_CONDENSING_CUIPM" : "PARTIAL_CONDENSING_HPIPM");
    if (!c.config || !c.config->qp_solver) { printf("create_from_name failed\n"); exit(2); }
#else
    ocp_qp_solver_plan_t plan; plan.qp_solver = PARTIAL_CONDENSING_HPIPM;
    c.config = ocp_qp_xcond_solver_config_create(plan);
    if (use_cuipm) ocp_qp_cuipm_config_initialize_default(c.config->qp_solver);
#endif
    c.dims = ocp_qp_xcond_solver_dims_create(c.config, NN);
    for (int k = 0; k <= NN; k++)
    {
        int nx = NX, nu = k < NN ? NU : 0, nbx = NX, nbu = nu, z = 0;
        c.config->dims_set(c.config, c.dims, k, "nx", &nx);
        c.config->dims_set(c.config, c.dims, k, "nu", &nu);
        c.config->dims_set(c.config, c.dims, k, "nbx", &nbx);
        c.config->dims_set(c.config, c.dims, k, "nbu", &nbu);
        c.config->dims_set(c.config, c.dims, k, "ng", &z);
        c.config->dims_set(c.config, c.dims, k, "ns", &z);
    }
    int nbxe0 = NX;
    c.config->dims_set(c.config, c.dims, 0, "nbxe", &nbxe0);
    c.opts = ocp_qp_xcond_solver_opts_create(c.config, c.dims);
    ocp_qp_xcond_solver_opts_set(c.config, c.opts, "cond_N", &N2);
    c.solver = ocp_qp_create(c.config, c.dims, c.opts);
    return c;
}

static ocp_qp_in *make_qp(chain *c, const double *x0)
{
    double A[NX * NX], B[NX * NU], b[NX], Q[NX * NX], R[NU * NU], S[NU * NX], q[NX], r[NU];
    mass_spring(0.5, A, B);
    memset(Q, 0, sizeof(Q)); memset(R, 0, sizeof(R)); memset(S, 0, sizeof(S));
    for (int i = 0; i < NX; i++) { Q[i + NX * i] = 1.0; q[i] = 0.1; b[i] = 0.1; }
    for (int i = 0; i < NU; i++) { R[i + NU * i] = 2.0; r[i] = 0.2; }
    ocp_qp_in *in = ocp_qp_in_create(c->dims->orig_dims);
    int idxb[NU + NX], idxe[NX];
    double lb[NU + NX], ub[NU + NX];
    for (int k = 0; k <= NN; k++)
    {
        int nu = k < NN ? NU : 0;
        if (k < NN) { d_ocp_qp_set_A(k, A, in); d_ocp_qp_set_B(k, B, in); d_ocp_qp_set_b(k, b, in); d_ocp_qp_set_R(k, R, in); d_ocp_qp_set_S(k, S, in); d_ocp_qp_set_r(k, r, in); }
        d_ocp_qp_set_Q(k, Q, in); d_ocp_qp_set_q(k, q, in);
        for (int i = 0; i < nu; i++) { idxb[i] = i; lb[i] = -0.5; ub[i] = 0.5; }
        for (int i = 0; i < NX; i++)
        {
            idxb[nu + i] = nu + i;
            lb[nu + i] = k == 0 ? x0[i] : -4.0;
            ub[nu + i] = k == 0 ? x0[i] : 4.0;
        }
        d_ocp_qp_set_idxb(k, idxb, in); d_ocp_qp_set_lb(k, lb, in); d_ocp_qp_set_ub(k, ub, in);
    }
    for (int i = 0; i < NX; i++) idxe[i] = i;
    d_ocp_qp_set_idxbxe(0, idxe, in);
    return in;
}

static double diff_out(ocp_qp_dims *d, ocp_qp_out *a, ocp_qp_out *b, double *du)
{
    double m = 0.0; *du = 0.0;
    for (int k = 0; k <= d->N; k++)
    {
        int n = d->nu[k] + d->nx[k], nc = 2 * (d->nb[k] + d->ng[k]);
        for (int i = 0; i < n; i++)
        {
            double e = fabs(BLASFEO_DVECEL(a->ux + k, i) - BLASFEO_DVECEL(b->ux + k, i));
            if (e > m) m = e;
            if (i < d->nu[k] && e > *du) *du = e;
        }
        if (k < d->N) for (int i = 0; i < d->nx[k + 1]; i++) { double e = fabs(BLASFEO_DVECEL(a->pi + k, i) - BLASFEO_DVECEL(b->pi + k, i)); if (e > m) m = e; }
        for (int i = 0; i < nc; i++)
        {
            double e = fabs(BLASFEO_DVECEL(a->lam + k, i) - BLASFEO_DVECEL(b->lam + k, i)); if (e > m) m = e;
            e = fabs(BLASFEO_DVECEL(a->t + k, i) - BLASFEO_DVECEL(b->t + k, i)); if (e > m) m = e;
        }
    }
    return m;
}

int main(void)
{
    int fails = 0;
    setvbuf(stdout, NULL, _IONBF, 0);
    double x0[NX] = {2.5, 2.5, 0, 0, 0, 0, 0, 0};
    int N2s[3] = {NN, 5, 3};
    for (int t = 0; t < 3; t++)
    {
        chain h = make_chain(0, N2s[t]), g = make_chain(1, N2s[t]);
        ocp_qp_in *in_h = make_qp(&h, x0), *in_g = make_qp(&g, x0);
        ocp_qp_out *out_h = ocp_qp_out_create(h.dims->orig_dims), *out_g = ocp_qp_out_create(g.dims->orig_dims);
        int st_h = ocp_qp_solve(h.solver, in_h, out_h), st_g = ocp_qp_solve(g.solver, in_g, out_g);
        /* reference acceptance test (test_qpsolvers.cpp:238-251): status 0 and residuals <= 1e-8 */
        double res[4];
        ocp_qp_inf_norm_residuals(g.dims->orig_dims, in_g, out_g, res);
        double du, dall = diff_out(h.dims->orig_dims, out_h, out_g, &du);
        int it_h, it_g;
        h.config->qp_solver->memory_get(h.config->qp_solver, ((ocp_qp_xcond_solver_memory *) h.solver->mem)->solver_memory, "iter", &it_h);
        g.config->qp_solver->memory_get(g.config->qp_solver, ((ocp_qp_xcond_solver_memory *) g.solver->mem)->solver_memory, "iter", &it_g);
        double maxres = fmax(fmax(res[0], res[1]), fmax(res[2], res[3]));
        int ok = st_h == 0 && st_g == 0 && it_h == it_g && du <= 1e-10 && dall <= 1e-6 && maxres <= 1e-8;
        printf("N2=%2d: status hpipm %d cuipm %d  iters %d %d  |du| %.2e  |dsol| %.2e  max residual (cuipm) %.2e  %s\n", N2s[t], st_h, st_g,
               it_h, it_g, du, dall, maxres, ok ? "OK" : "FAIL");
        fails += !ok;
        if (t == 0)
        {
            /* Riccati getters through the xcond solver (ocp_qp_xcond_solver.c:424) */
            double Kh[NU * NX], Kg[NU * NX], Ph[NX * NX], Pg[NX * NX], e = 0.0;
            h.config->solver_get(h.config, in_h, out_h, h.opts, h.solver->mem, "K", 2, Kh, NU, NX);
            g.config->solver_get(g.config, in_g, out_g, g.opts, g.solver->mem, "K", 2, Kg, NU, NX);
            h.config->solver_get(h.config, in_h, out_h, h.opts, h.solver->mem, "P", 2, Ph, NX, NX);
            g.config->solver_get(g.config, in_g, out_g, g.opts, g.solver->mem, "P", 2, Pg, NX, NX);
            for (int i = 0; i < NU * NX; i++) e = fmax(e, fabs(Kh[i] - Kg[i]));
            for (int i = 0; i < NX * NX; i++) e = fmax(e, fabs(Ph[i] - Pg[i]));
            printf("getters K,P at stage 2: max diff %.2e %s\n", e, e <= 1e-6 ? "OK" : "FAIL");
            fails += !(e <= 1e-6);
            /* solution sensitivities through the xcond solver (ocp_qp_xcond_solver.c:672-726): same seed, both plugins */
            for (int adj = 0; adj < 2; adj++)
            {
                ocp_qp_dims *od = h.dims->orig_dims;
                void *sm_h = calloc(1, ocp_qp_seed_calculate_size(od) + 64), *sm_g = calloc(1, ocp_qp_seed_calculate_size(od) + 64);
                ocp_qp_seed *sd_h = ocp_qp_seed_assign(od, sm_h), *sd_g = ocp_qp_seed_assign(od, sm_g);
                ocp_qp_out *se_h = ocp_qp_out_create(od), *se_g = ocp_qp_out_create(od);
                for (int k = 0; k <= od->N; k++)
                {
                    int n = od->nu[k] + od->nx[k], nc = 2 * (od->nb[k] + od->ng[k]);
                    for (int i = 0; i < n; i++) { double v = sin(1.0 + i + 3.0 * k); BLASFEO_DVECEL(sd_h->seed_g + k, i) = v; BLASFEO_DVECEL(sd_g->seed_g + k, i) = v; }
                    if (k < od->N) for (int i = 0; i < od->nx[k + 1]; i++) { double v = 0.1 * cos(2.0 + i + k); BLASFEO_DVECEL(sd_h->seed_b + k, i) = v; BLASFEO_DVECEL(sd_g->seed_b + k, i) = v; }
                    for (int i = 0; i < nc; i++) { BLASFEO_DVECEL(sd_h->seed_d + k, i) = 0.0; BLASFEO_DVECEL(sd_g->seed_d + k, i) = 0.0; BLASFEO_DVECEL(sd_h->seed_m + k, i) = 0.0; BLASFEO_DVECEL(sd_g->seed_m + k, i) = 0.0; }
                }
                if (adj)
                {
                    h.config->eval_adj_sens(h.config, h.dims, in_h, sd_h, se_h, h.opts, h.solver->mem, h.solver->work);
                    g.config->eval_adj_sens(g.config, g.dims, in_g, sd_g, se_g, g.opts, g.solver->mem, g.solver->work);
                }
                else
                {
                    h.config->eval_forw_sens(h.config, h.dims, in_h, sd_h, se_h, h.opts, h.solver->mem, h.solver->work);
                    g.config->eval_forw_sens(g.config, g.dims, in_g, sd_g, se_g, g.opts, g.solver->mem, g.solver->work);
                }
                double es = 0.0, sc = 0.0;
                for (int k = 0; k <= od->N; k++)
                    for (int i = 0; i < od->nu[k] + od->nx[k]; i++)
                    {
                        es = fmax(es, fabs(BLASFEO_DVECEL(se_h->ux + k, i) - BLASFEO_DVECEL(se_g->ux + k, i)));
                        sc = fmax(sc, fabs(BLASFEO_DVECEL(se_h->ux + k, i)));
                    }
                printf("%s sensitivities (dux): max diff %.2e of %.2e %s\n", adj ? "adjoint" : "forward", es, sc, es <= 1e-6 * sc && sc > 0 ? "OK" : "FAIL");
                fails += !(es <= 1e-6 * sc && sc > 0);
                ocp_qp_out_free(se_h); ocp_qp_out_free(se_g); free(sm_h); free(sm_g);
            }
            double tq; int sm;
            g.config->qp_solver->memory_get(g.config->qp_solver, ((ocp_qp_xcond_solver_memory *) g.solver->mem)->solver_memory, "time_qp_solver_call", &tq);
            g.config->qp_solver->memory_get(g.config->qp_solver, ((ocp_qp_xcond_solver_memory *) g.solver->mem)->solver_memory, "stat_m", &sm);
            fails += !(tq > 0 && sm == CUIPM_STAT_M);
        }
        g.config->terminate(g.config, g.solver->mem, g.solver->work);
        h.config->terminate(h.config, h.solver->mem, h.solver->work);
    }
    {
        /* batched entry at the plugin level: QPs with x0 already eliminated (stage 0: nx = 0, b_0 = b + A x0), solved
         * one by one with the reference's HPIPM plugin and in one launch with ocp_qp_cuipm_batch_solve */
        enum { NB = 6 };
        double A[NX * NX], B[NX * NU], Q[NX * NX], R[NU * NU], q[NX], r[NU], lb[NU + NX], ub[NU + NX];
        int idxb[NU + NX];
        mass_spring(0.5, A, B);
        memset(Q, 0, sizeof(Q)); memset(R, 0, sizeof(R));
        for (int i = 0; i < NX; i++) { Q[i + NX * i] = 1.0; q[i] = 0.1; }
        for (int i = 0; i < NU; i++) { R[i + NU * i] = 2.0; r[i] = 0.2; }
        ocp_qp_dims *dims = ocp_qp_dims_create(NN);
        for (int k = 0; k <= NN; k++)
        {
            d_ocp_qp_dim_set_nx(k, k == 0 ? 0 : NX, dims);
            d_ocp_qp_dim_set_nu(k, k < NN ? NU : 0, dims);
            d_ocp_qp_dim_set_nbx(k, k == 0 ? 0 : NX, dims);
            d_ocp_qp_dim_set_nbu(k, k < NN ? NU : 0, dims);
        }
        qp_solver_config cfg_h, cfg_g;
        ocp_qp_hpipm_config_initialize_default(&cfg_h);
        ocp_qp_cuipm_config_initialize_default(&cfg_g);
        void *oh = cfg_h.opts_assign(&cfg_h, dims, calloc(1, cfg_h.opts_calculate_size(&cfg_h, dims)));
        void *og = cfg_g.opts_assign(&cfg_g, dims, calloc(1, cfg_g.opts_calculate_size(&cfg_g, dims)));
        cfg_h.opts_initialize_default(&cfg_h, dims, oh); cfg_g.opts_initialize_default(&cfg_g, dims, og);
        void *mh = cfg_h.memory_assign(&cfg_h, dims, oh, calloc(1, cfg_h.memory_calculate_size(&cfg_h, dims, oh)));
        void *mg = cfg_g.memory_assign(&cfg_g, dims, og, calloc(1, cfg_g.memory_calculate_size(&cfg_g, dims, og)));
        ocp_qp_in *ins[NB];
        ocp_qp_out *outs[NB], *ref[NB];
        int status[NB];
        double worst = 0.0;
        for (int i = 0; i < NB; i++)
        {
            double x0i[NX], b0[NX], bb[NX];
            for (int j = 0; j < NX; j++) { x0i[j] = x0[j] + 0.3 * sin(1.0 + i + 2.0 * j); bb[j] = 0.1; }
            for (int j = 0; j < NX; j++) { b0[j] = 0.1; for (int c = 0; c < NX; c++) b0[j] += A[j + NX * c] * x0i[c]; }
            ins[i] = ocp_qp_in_create(dims);
            for (int k = 0; k <= NN; k++)
            {
                int nu = k < NN ? NU : 0, nx = k == 0 ? 0 : NX;
                if (k < NN)
                {
                    if (k > 0) d_ocp_qp_set_A(k, A, ins[i]);
                    d_ocp_qp_set_B(k, B, ins[i]); d_ocp_qp_set_b(k, k == 0 ? b0 : bb, ins[i]);
                    d_ocp_qp_set_R(k, R, ins[i]); d_ocp_qp_set_r(k, r, ins[i]);
                }
                if (k > 0) { d_ocp_qp_set_Q(k, Q, ins[i]); d_ocp_qp_set_q(k, q, ins[i]); }
                for (int j = 0; j < nu; j++) { idxb[j] = j; lb[j] = -0.5; ub[j] = 0.5; }
                for (int j = 0; j < nx; j++) { idxb[nu + j] = nu + j; lb[nu + j] = -4.0; ub[nu + j] = 4.0; }
                d_ocp_qp_set_idxb(k, idxb, ins[i]); d_ocp_qp_set_lb(k, lb, ins[i]); d_ocp_qp_set_ub(k, ub, ins[i]);
            }
            outs[i] = ocp_qp_out_create(dims);
            ref[i] = ocp_qp_out_create(dims);
            cfg_h.evaluate(&cfg_h, ins[i], ref[i], oh, mh, NULL);
        }
        int rc = ocp_qp_cuipm_batch_solve(&cfg_g, NB, ins, outs, og, mg, status);
        for (int i = 0; i < NB; i++)
        {
            double du, dall = diff_out(dims, ref[i], outs[i], &du);
            worst = fmax(worst, du);
            fails += status[i] != 0 || dall > 1e-6;
        }
        printf("batch entry: %d QPs, rc %d, max |du| vs per-instance HPIPM plugin %.2e %s\n", NB, rc, worst, (rc == 0 && worst <= 1e-10) ? "OK" : "FAIL");
        fails += !(rc == 0 && worst <= 1e-10);
        cfg_g.terminate(&cfg_g, mg, NULL);
    }
    {
        /* the xcond chain as a batched entry: uncondensed QPs with x0 as stage-0 equality bounds (as ocp_qp_xcond_solver holds
         * them), elimination + block condensing + solve + expansion + restore on the device, against the reference chain
         * (PARTIAL_CONDENSING_HPIPM behind ocp_qp_xcond_solver with acados' CPU condensing) instance by instance; then the
         * SQP-RTI split: condense_lhs on the batch, new x0 (vectors only), condense_rhs_and_solve */
        enum { NBX = 5 };
        int conds[2] = {NN, 5};
        for (int t = 0; t < 2; t++)
        {
            chain h = make_chain(0, conds[t]);
            ocp_qp_in *ins[NBX], *ins2[NBX];
            ocp_qp_out *outs[NBX], *ref[NBX], *ref2[NBX];
            int status[NBX];
            for (int i = 0; i < NBX; i++)
            {
                double xa[NX], xb[NX];
                for (int j = 0; j < NX; j++) { xa[j] = x0[j] + 0.3 * sin(1.0 + i + 2.0 * j); xb[j] = xa[j] + 0.1 * cos(3.0 + i + j); }
                ins[i] = make_qp(&h, xa); ins2[i] = make_qp(&h, xb);
                outs[i] = ocp_qp_out_create(h.dims->orig_dims); ref[i] = ocp_qp_out_create(h.dims->orig_dims); ref2[i] = ocp_qp_out_create(h.dims->orig_dims);
                fails += ocp_qp_solve(h.solver, ins[i], ref[i]) != 0;
                fails += ocp_qp_solve(h.solver, ins2[i], ref2[i]) != 0;
            }
            qp_solver_config cfg_g;
            ocp_qp_cuipm_config_initialize_default(&cfg_g);
            ocp_qp_dims *od = h.dims->orig_dims;
            void *og = cfg_g.opts_assign(&cfg_g, od, calloc(1, cfg_g.opts_calculate_size(&cfg_g, od)));
            cfg_g.opts_initialize_default(&cfg_g, od, og);
            ocp_qp_cuipm_xcond_batch *ctx = ocp_qp_cuipm_xcond_batch_create(ins[0], NBX, conds[t], 0);
            if (!ctx) { printf("xcond batch entry: create failed FAIL\n"); fails++; continue; }
            int rc = ocp_qp_cuipm_xcond_batch_solve(ctx, NBX, ins, outs, og, 0, status);
            double worst = 0.0, wall = 0.0;
            for (int i = 0; i < NBX; i++)
            {
                double du, dall = diff_out(od, ref[i], outs[i], &du);
                worst = fmax(worst, du); wall = fmax(wall, dall);
                fails += status[i] != 0;
            }
            int ok = rc == 0 && worst <= 1e-10 && wall <= 1e-6;
            printf("xcond batch entry N2=%2d: %d QPs, rc %d, max |du| %.2e |dsol| %.2e vs the reference chain %s\n", conds[t], NBX, rc, worst, wall, ok ? "OK" : "FAIL");
            fails += !ok;
            /* RTI split */
            rc = ocp_qp_cuipm_xcond_batch_solve(ctx, NBX, ins, outs, og, 1, status);
            rc |= ocp_qp_cuipm_xcond_batch_solve(ctx, NBX, ins2, outs, og, 2, status);
            worst = 0.0; wall = 0.0;
            for (int i = 0; i < NBX; i++)
            {
                double du, dall = diff_out(od, ref2[i], outs[i], &du);
                worst = fmax(worst, du); wall = fmax(wall, dall);
                fails += status[i] != 0;
            }
            ok = rc == 0 && worst <= 1e-10 && wall <= 1e-6;
            printf("xcond batch entry N2=%2d, condense_lhs then condense_rhs_and_solve with new x0: max |du| %.2e |dsol| %.2e %s\n", conds[t], worst, wall, ok ? "OK" : "FAIL");
            fails += !ok;
            ocp_qp_cuipm_xcond_batch_destroy(ctx);
            h.config->terminate(h.config, h.solver->mem, h.solver->work);
        }
    }
    printf(fails ? "PLUGIN TEST FAILED (%d)\n" : "PLUGIN TEST PASSED\n", fails);
    return fails != 0;
}
