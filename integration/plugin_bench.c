#define _POSIX_C_SOURCE 199309L
/*
 * plugin_bench.c -- harness (our code) around the plugin's batched entry point, linked against the patched libacados of
 * integration/Makefile (_build/libacados_cuipm.so: reference objects + plugin + libcuipm).
 *
 * A caller inside acados holds n `ocp_qp_in` / `ocp_qp_out` objects (HPIPM's panel-major structs, one contiguous
 * allocation each, interfaces/acados_c/ocp_qp_interface.c:378-386); ocp_qp_cuipm_batch_solve() is what it would call instead
 * of the reference's OpenMP loop over per-instance evaluate() (c_templates_tera/acados_solver.in.c:3223-3243).  This file
 * builds those objects from QP records (the benchmark's workload), calls the entry point `reps` times and reports the
 * wall-clock time per call -- struct unpacking on the host threads, page-locked staging, H2D, kernels, D2H, struct packing:
 * everything a user of the plugin pays.  bench.py reports it as e2e_plugin; tests/test_plugin.py checks the solutions.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "acados/ocp_qp/ocp_qp_common.h"
#include "acados_c/ocp_qp_interface.h"
#include "blasfeo/include/blasfeo_d_aux.h"
#include "hpipm/include/hpipm_d_ocp_qp.h"
#include "hpipm/include/hpipm_d_ocp_qp_dim.h"

#include "cuipm.h"
#include "ocp_qp_cuipm.h"

static double now_s(void)
{
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec + 1e-9 * ts.tv_nsec;
}

typedef struct
{
    int n;
    cuipm_layout *l;
    qp_solver_config *config;
    ocp_qp_dims *dims;
    ocp_qp_in **in;
    ocp_qp_out **out;
    void *opts_mem, *mem_mem, *opts, *mem;
    int *status;
} plugin_bench;

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

static void load_qp(ocp_qp_in *in, const cuipm_shape *sh, const cuipm_layout *l, const double *q)
{
    for (int k = 0; k <= sh->N; k++)
    {
        int n = sh->nu[k] + sh->nx[k], nc = 2 * (sh->nb[k] + sh->ng[k] + sh->ns[k]);
        if (sh->nb[k] > 0) d_ocp_qp_set_idxb(k, (int *) sh->idxb[k], in);
        if (sh->nb[k] + sh->ng[k] > 0 && sh->ns[k] > 0) d_ocp_qp_set_idxs_rev(k, (int *) sh->idxs_rev[k], in);
        if (k < sh->N)
        {
            blasfeo_pack_dmat(n, sh->nx[k + 1], (double *) q + l->off_BAt[k], n, in->BAbt + k, 0, 0);
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

/* n QPs of shape sh from the records qp (cuipm layout); opts: name/value pairs are taken from co.  Returns NULL on failure. */
void *plugin_bench_create(const cuipm_shape *sh, int n, const double *qp, const cuipm_opts *co)
{
    plugin_bench *b = (plugin_bench *) calloc(1, sizeof(plugin_bench));
    b->n = n;
    b->l = cuipm_layout_create(sh);
    b->config = (qp_solver_config *) calloc(1, sizeof(qp_solver_config));
    ocp_qp_cuipm_config_initialize_default(b->config);
    b->dims = make_dims(sh);
    b->opts_mem = calloc(1, b->config->opts_calculate_size(b->config, b->dims));
    b->opts = b->config->opts_assign(b->config, b->dims, b->opts_mem);
    b->config->opts_initialize_default(b->config, b->dims, b->opts);
    ((ocp_qp_cuipm_opts *) b->opts)->c = *co;
    b->config->opts_update(b->config, b->dims, b->opts);
    b->mem_mem = calloc(1, b->config->memory_calculate_size(b->config, b->dims, b->opts));
    b->mem = b->config->memory_assign(b->config, b->dims, b->opts, b->mem_mem);
    b->in = (ocp_qp_in **) calloc((size_t) n, sizeof(ocp_qp_in *));
    b->out = (ocp_qp_out **) calloc((size_t) n, sizeof(ocp_qp_out *));
    b->status = (int *) calloc((size_t) n, sizeof(int));
    for (int i = 0; i < n; i++)
    {
        b->in[i] = ocp_qp_in_create(b->dims);
        b->out[i] = ocp_qp_out_create(b->dims);
        load_qp(b->in[i], sh, b->l, qp + b->l->qp_stride * (size_t) i);
    }
    return b;
}

/* `reps` calls of ocp_qp_cuipm_batch_solve over all n QPs; seconds[r] = wall-clock time of call r.  Returns the worst status. */
int plugin_bench_run(void *b_, int reps, double *seconds)
{
    plugin_bench *b = (plugin_bench *) b_;
    int worst = 0;
    for (int r = 0; r < reps; r++)
    {
        const double t0 = now_s();
        int st = ocp_qp_cuipm_batch_solve(b->config, b->n, b->in, b->out, b->opts, b->mem, b->status);
        seconds[r] = now_s() - t0;
        if (st != 0) worst = st;
    }
    return worst;
}

/* solutions of the last call as solution records + per-QP iteration counts / acados status */
void plugin_bench_get(void *b_, const cuipm_shape *sh, double *sol, int *iter, int *status)
{
    plugin_bench *b = (plugin_bench *) b_;
    for (int i = 0; i < b->n; i++)
    {
        double *s = sol + b->l->sol_stride * (size_t) i;
        ocp_qp_out *o = b->out[i];
        for (int k = 0; k <= sh->N; k++)
        {
            int n = sh->nu[k] + sh->nx[k], nc = 2 * (sh->nb[k] + sh->ng[k] + sh->ns[k]);
            blasfeo_unpack_dvec(n + 2 * sh->ns[k], o->ux + k, 0, s + b->l->off_ux[k], 1);
            if (k < sh->N) blasfeo_unpack_dvec(sh->nx[k + 1], o->pi + k, 0, s + b->l->off_pi[k], 1);
            blasfeo_unpack_dvec(nc, o->lam + k, 0, s + b->l->off_lam[k], 1);
            blasfeo_unpack_dvec(nc, o->t + k, 0, s + b->l->off_t[k], 1);
        }
        iter[i] = ((qp_info *) o->misc)->num_iter;
        status[i] = b->status[i];
    }
}

void plugin_bench_destroy(void *b_)
{
    plugin_bench *b = (plugin_bench *) b_;
    b->config->terminate(b->config, b->mem, NULL);
    for (int i = 0; i < b->n; i++) { ocp_qp_in_free(b->in[i]); ocp_qp_out_free(b->out[i]); }
    free(b->in); free(b->out); free(b->status); free(b->mem_mem); free(b->opts_mem);
    ocp_qp_dims_free(b->dims); free(b->config); cuipm_layout_destroy(b->l); free(b);
}
