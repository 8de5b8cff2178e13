// cuipm_host.cpp -- host-only part of the C ABI (include/cuipm.h): option handling and record layout.
// No CUDA here, so these entry points work on a machine without a GPU (the solve entry points do not).
//
// Option defaults restate OCP_QP_IPM_ARG_SET_DEFAULT (external/hpipm/ocp_qp/x_ocp_qp_ipm.c:69-260) and the
// overrides acados applies after every mode switch (acados/ocp_qp/ocp_qp_hpipm.c:101-113).
#include <cstdlib>
#include <cstring>
#include <string>

#include "cuipm.h"
#include "cuipm_internal.h"

namespace cuipm {
thread_local std::string g_last_error;
void set_error(const std::string &msg) { g_last_error = msg; }
}  // namespace cuipm

extern "C" const char *cuipm_last_error(void) { return cuipm::g_last_error.c_str(); }

extern "C" void cuipm_opts_set_default(cuipm_opts *o, int mode)
{
    std::memset(o, 0, sizeof(*o));
    o->mode = mode;
    // common to all modes
    o->alpha_min = 1e-12;
    o->res_g_max = 1e-6;
    o->res_b_max = 1e-8;
    o->res_d_max = 1e-8;
    o->res_m_max = 1e-8;
    o->dual_gap_max = 1e15;
    o->pred_corr = 1;
    o->cond_pred_corr = 1;
    o->itref_pred_max = 0;
    o->reg_prim = 1e-15;
    o->lam_min = 1e-16;
    o->t_min = 1e-16;
    o->tau_min = 1e-16;
    o->lam0_min = 1e-9;
    o->t0_min = 1e-9;
    o->warm_start = 0;
    o->abs_form = 0;
    o->comp_dual_sol_eq = 1;
    o->comp_res_exit = 1;
    o->var_init_scheme = 0;
    o->t_lam_min = 2;
    o->t0_init = 2;
    o->m_relax = 0.0;
    switch (mode)
    {
        case CUIPM_SPEED_ABS:
            o->mu0 = 1e1; o->res_g_max = 1e0; o->res_b_max = 1e0; o->res_d_max = 1e0; o->iter_max = 15;
            o->itref_corr_max = 0; o->lq_fact = 0; o->abs_form = 1; o->comp_dual_sol_eq = 0; o->comp_res_exit = 0;
            o->split_step = 1;
            break;
        case CUIPM_SPEED:
            o->mu0 = 1e1; o->iter_max = 15; o->itref_corr_max = 0; o->lq_fact = 0; o->split_step = 1;
            break;
        case CUIPM_ROBUST:
            o->mu0 = 1e2; o->iter_max = 100; o->itref_corr_max = 4; o->lq_fact = 2; o->split_step = 0;
            break;
        case CUIPM_BALANCE:
        default:
            o->mode = CUIPM_BALANCE;
            o->mu0 = 1e1; o->iter_max = 30; o->itref_corr_max = 2; o->lq_fact = 1; o->split_step = 0;
            break;
    }
    o->stat_max = o->iter_max;
}

extern "C" void cuipm_opts_set_default_acados(cuipm_opts *o, int mode)
{
    cuipm_opts_set_default(o, mode);
    // ocp_qp_hpipm_opts_overwrite_mode_opts, acados/ocp_qp/ocp_qp_hpipm.c:101-113
    o->res_g_max = 1e-6;
    o->res_b_max = 1e-8;
    o->res_d_max = 1e-8;
    o->res_m_max = 1e-8;
    o->iter_max = 50;
    o->stat_max = 50;
    o->alpha_min = 1e-8;
    o->mu0 = 1e0;
    o->var_init_scheme = 1;
}

namespace {
enum FieldType { F_INT, F_DBL };
struct Field { const char *name; FieldType type; size_t off; };
#define FI(n, m) {n, F_INT, offsetof(cuipm_opts, m)}
#define FD(n, m) {n, F_DBL, offsetof(cuipm_opts, m)}
const Field kFields[] = {
    FI("iter_max", iter_max), FD("alpha_min", alpha_min), FD("mu0", mu0), FD("tol_stat", res_g_max),
    FD("tol_eq", res_b_max), FD("tol_ineq", res_d_max), FD("tol_comp", res_m_max), FD("tol_dual_gap", dual_gap_max),
    FD("reg_prim", reg_prim), FI("warm_start", warm_start), FI("pred_corr", pred_corr),
    FI("cond_pred_corr", cond_pred_corr), FI("comp_dual_sol_eq", comp_dual_sol_eq), FI("comp_res_exit", comp_res_exit),
    FD("lam_min", lam_min), FD("t_min", t_min), FD("tau_min", tau_min), FD("lam0_min", lam0_min), FD("t0_min", t0_min),
    FI("split_step", split_step), FI("var_init_scheme", var_init_scheme), FI("t_lam_min", t_lam_min),
    FI("t0_init", t0_init), FI("itref_corr_max", itref_corr_max), FI("itref_pred_max", itref_pred_max),
    FI("lq_fact", lq_fact), FI("stat_max", stat_max), FD("m_relax", m_relax),
};
const Field *find_field(const char *name)
{
    for (const Field &f : kFields)
        if (!std::strcmp(f.name, name)) return &f;
    return nullptr;
}
}  // namespace

extern "C" int cuipm_opts_set(cuipm_opts *o, const char *field, const void *value)
{
    if (!std::strcmp(field, "hpipm_mode"))
    {
        const char *m = (const char *) value;
        int mode;
        if (!std::strcmp(m, "BALANCE")) mode = CUIPM_BALANCE;
        else if (!std::strcmp(m, "SPEED")) mode = CUIPM_SPEED;
        else if (!std::strcmp(m, "SPEED_ABS")) mode = CUIPM_SPEED_ABS;
        else if (!std::strcmp(m, "ROBUST")) mode = CUIPM_ROBUST;
        else return CUIPM_ERR_INVALID;
        cuipm_opts_set_default_acados(o, mode);
        return CUIPM_OK;
    }
    if (!std::strcmp(field, "ric_alg"))
    {   // only the square-root Riccati algorithm exists here (HPIPM's default, square_root_alg=1)
        return *(const int *) value == 1 ? CUIPM_OK : CUIPM_ERR_INVALID;
    }
    if (!std::strcmp(field, "comp_res_pred") || !std::strcmp(field, "update_fact_exit") || !std::strcmp(field, "m_safe"))
        return CUIPM_OK;  // accepted, no effect on this path (split_step=0, factorisation always current)
    const Field *f = find_field(field);
    if (!f) return CUIPM_ERR_INVALID;
    if (f->type == F_INT) *(int *) ((char *) o + f->off) = *(const int *) value;
    else *(double *) ((char *) o + f->off) = *(const double *) value;
    if (!std::strcmp(field, "iter_max") && o->stat_max < o->iter_max) o->stat_max = o->iter_max;
    return CUIPM_OK;
}

extern "C" int cuipm_opts_get(const cuipm_opts *o, const char *field, void *value)
{
    const Field *f = find_field(field);
    if (!f) return CUIPM_ERR_INVALID;
    if (f->type == F_INT) *(int *) value = *(const int *) ((const char *) o + f->off);
    else *(double *) value = *(const double *) ((const char *) o + f->off);
    return CUIPM_OK;
}

int cuipm::opts_check(const cuipm_opts *o)
{
    if (o->abs_form != 0) { set_error("abs_form=1 (SPEED_ABS) is not supported"); return CUIPM_ERR_INVALID; }
    if (o->split_step != 0) { set_error("split_step=1 (SPEED modes) is not supported"); return CUIPM_ERR_INVALID; }
    if (o->comp_dual_sol_eq != 1 || o->comp_res_exit != 1) { set_error("comp_dual_sol_eq/comp_res_exit must be 1"); return CUIPM_ERR_INVALID; }
    if (o->var_init_scheme != 1) { set_error("only var_init_scheme=1 (the acados default) is supported"); return CUIPM_ERR_INVALID; }
    // m != 0 changes more than the complementarity residual: COMPUTE_MU_AFF_QP subtracts m and COMPUTE_ALPHA_QP switches to a
    // SEQUENTIAL ratio test with a quadratic root per violated constraint (x_core_qp_ipm_aux.c:397-436), which the fused,
    // parallel ratio test of the kernels does not reproduce
    if (o->m_relax < 0.0) { set_error("tau_min / m relaxation must be >= 0"); return CUIPM_ERR_INVALID; }
    if (o->itref_pred_max != 0) { set_error("itref_pred_max must be 0"); return CUIPM_ERR_INVALID; }
    // stat_max may be smaller than iter_max: the kernels write row kk+1 of the statistics table only while kk+1 < stat_max
    if (o->iter_max < 0 || o->stat_max < 0) { set_error("need iter_max >= 0 and stat_max >= 0"); return CUIPM_ERR_INVALID; }
    if (o->itref_corr_max < 0 || o->itref_corr_max > 8) { set_error("itref_corr_max out of range"); return CUIPM_ERR_INVALID; }
    return CUIPM_OK;
}

// ---- layout ----------------------------------------------------------------------------------------

static inline size_t ev2(size_t n) { return (n + 1) & ~(size_t) 1; }

extern "C" cuipm_layout *cuipm_layout_create(const cuipm_shape *sh)
{
    if (!sh || sh->N < 0) return nullptr;
    const int N = sh->N;
    cuipm_layout *l = (cuipm_layout *) std::calloc(1, sizeof(cuipm_layout));
    l->N = N;
    size_t **arrs[] = {&l->qp_stage, &l->off_BAt, &l->off_RSQ, &l->off_DCt, &l->off_b, &l->off_rq, &l->off_d,
                       &l->off_dmask, &l->off_Z, &l->off_z, &l->sol_stage, &l->off_ux, &l->off_pi, &l->off_lam,
                       &l->off_t};
    for (size_t **a : arrs) *a = (size_t *) std::calloc((size_t) N + 2, sizeof(size_t));
    size_t q = 0, s = 0;
    for (int k = 0; k <= N; k++)
    {
        const size_t n = (size_t) sh->nu[k] + sh->nx[k];
        const size_t nx1 = k < N ? (size_t) sh->nx[k + 1] : 0;
        const size_t nc = 2 * ((size_t) sh->nb[k] + sh->ng[k] + sh->ns[k]);
        const size_t ns2 = 2 * (size_t) sh->ns[k];
        l->qp_stage[k] = q;
        l->off_BAt[k] = q;   q += ev2(n * nx1);
        l->off_RSQ[k] = q;   q += ev2(n * n);
        l->off_DCt[k] = q;   q += ev2(n * (size_t) sh->ng[k]);
        l->off_b[k] = q;     q += ev2(nx1);
        l->off_rq[k] = q;    q += ev2(n);
        l->off_d[k] = q;     q += ev2(nc);
        l->off_dmask[k] = q; q += ev2(nc);
        l->off_Z[k] = q;     q += ev2(ns2);
        l->off_z[k] = q;     q += ev2(ns2);
        l->sol_stage[k] = s;
        l->off_ux[k] = s;    s += ev2(n + ns2);
        l->off_pi[k] = s;    s += ev2(nx1);
        l->off_lam[k] = s;   s += ev2(nc);
        l->off_t[k] = s;     s += ev2(nc);
    }
    l->qp_stage[N + 1] = q;
    l->sol_stage[N + 1] = s;
    l->qp_stride = q;
    l->sol_stride = s;
    return l;
}

extern "C" void cuipm_layout_destroy(cuipm_layout *l)
{
    if (!l) return;
    size_t *arrs[] = {l->qp_stage, l->off_BAt, l->off_RSQ, l->off_DCt, l->off_b, l->off_rq, l->off_d, l->off_dmask,
                      l->off_Z, l->off_z, l->sol_stage, l->off_ux, l->off_pi, l->off_lam, l->off_t};
    for (size_t *a : arrs) std::free(a);
    std::free(l);
}
