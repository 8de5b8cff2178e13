/*
 * oracle_ipm.h -- TEST INFRASTRUCTURE ONLY.
 *
 * Plain-C, single-threaded-per-QP restatement of the reference algorithm on the hot path
 * (HPIPM's OCP-QP interior-point method as configured by acados), operating on the same record
 * layout as the product's C ABI (include/cuipm.h) so that tests can hand both the same buffers.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load it.
 * The product library never links or calls anything in oracle/.
 */
#ifndef ORACLE_IPM_H_
#define ORACLE_IPM_H_

#include "../include/cuipm.h"   /* types only: cuipm_shape, cuipm_opts, cuipm_layout, cuipm_info */

#ifdef __cplusplus
extern "C" {
#endif

/* independent re-computation of the record layout (checked against cuipm_layout_create in tests) */
cuipm_layout *oracle_layout_create(const cuipm_shape *shape);
void oracle_layout_destroy(cuipm_layout *l);

/* Solve nbatch QPs on the CPU with `nthreads` OpenMP threads (<=0: all).  Same buffers as cuipm_solve_host. */
int oracle_solve(const cuipm_shape *shape, int nbatch, const double *qp, double *sol, cuipm_info *info,
                 double *stat, const cuipm_opts *opts, int nthreads);

/* oracle_solve followed by the solution sensitivities for one seed per QP (restates d_ocp_qp_ipm_sens_frw / _adj,
 * external/hpipm/ocp_qp/x_ocp_qp_ipm.c:3285-3444).  seed and sens are records in the solution layout:
 * (seed_g, seed_b, seed_d, seed_m) in the (ux, pi, lam, t) slots. */
int oracle_solve_sens(const cuipm_shape *shape, int nbatch, const double *qp, double *sol, cuipm_info *info,
                      const cuipm_opts *opts, int nthreads, const double *seed, double *sens, int adjoint);

/* KKT residuals of a given primal-dual point (restates OCP_QP_RES_COMPUTE + INF_NORM,
 * external/hpipm/ocp_qp/x_ocp_qp_res.c:345-531,689-727): res_max[4], mu, obj, dual_gap per QP. */
int oracle_residuals(const cuipm_shape *shape, int nbatch, const double *qp, const double *sol,
                     cuipm_info *info);

#ifdef __cplusplus
}
#endif
#endif
