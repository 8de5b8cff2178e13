// condense_emul.cpp -- TEST INFRASTRUCTURE ONLY.
//
// Runs the product's block-condensing code (acados_b200/csrc/cuipm_condense_core.h, the body of the CUDA kernels) on the host
// with a sequential execution policy: the threads of a phase one after the other, a phase boundary where the kernel has a
// barrier.  It lets the CPU test-suite check the kernel's arithmetic and index maps against acados_b200/condensing.py (which
// is pinned against the reference) without a GPU; the GPU tests then only have to confirm that the parallel execution agrees.
// The product never calls this.
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../acados_b200/csrc/cuipm_condense_plan.h"

using namespace cuipm_cond;

namespace {
struct SeqExec
{
    int nt;
    int nthreads() const { return nt; }
    template <class F> void phase(F f) { for (int t = 0; t < nt; t++) f(t); }
};
}  // namespace

extern "C" {

// sizes of the condensed records for buffer allocation; returns 0 on a bad cond_N
int emul_condensed_strides(const cuipm_shape *sh, int cond_N, size_t *qp_stride, size_t *sol_stride)
{
    HostPlan hp;
    if (!build_plan(sh, cond_N, hp)) return 0;
    *qp_stride = hp.lc->qp_stride; *sol_stride = hp.lc->sol_stride;
    return 1;
}

int emul_condense(const cuipm_shape *sh, int cond_N, int nbatch, const double *qp, double *qp2, int nthreads)
{
    HostPlan hp;
    if (!build_plan(sh, cond_N, hp)) return -1;
    const Plan P = hp.plan(hp.ipool.data(), hp.upool.data());
    std::vector<double> scr(scratch_doubles(P));
    SeqExec ex{nthreads};
    for (int q = 0; q < nbatch; q++)
    {
        double *o = qp2 + (size_t) q * hp.lc->qp_stride;
        condense_one<COND_ALL>(ex, P, qp + (size_t) q * hp.lo->qp_stride, o, scr.data(), nullptr);
    }
    return 0;
}

// the lhs / rhs split: lhs pass on the records qp (condensed records qp2 written in full, T_j kept in tbuf, t_stride doubles per
// QP as emul_t_stride says), then -- if qp_new is given -- the rhs pass on qp_new (same matrices, new vectors) refreshing qp2
size_t emul_t_stride(const cuipm_shape *sh, int cond_N)
{
    HostPlan hp;
    if (!build_plan(sh, cond_N, hp)) return 0;
    return hp.t_stride;
}

int emul_condense_split(const cuipm_shape *sh, int cond_N, int nbatch, const double *qp, const double *qp_new, double *qp2, int nthreads)
{
    HostPlan hp;
    if (!build_plan(sh, cond_N, hp)) return -1;
    const Plan P = hp.plan(hp.ipool.data(), hp.upool.data());
    std::vector<double> scr(scratch_doubles(P)), tb((size_t) hp.t_stride * nbatch);
    SeqExec ex{nthreads};
    for (int q = 0; q < nbatch; q++)
        condense_one<COND_LHS>(ex, P, qp + (size_t) q * hp.lo->qp_stride, qp2 + (size_t) q * hp.lc->qp_stride, scr.data(), tb.data() + (size_t) q * hp.t_stride);
    if (qp_new)
        for (int q = 0; q < nbatch; q++)
        {
            for (double &x : scr) x = 1e300;      // the rhs pass must not depend on what the lhs pass left in the scratch
            condense_one<COND_RHS>(ex, P, qp_new + (size_t) q * hp.lo->qp_stride, qp2 + (size_t) q * hp.lc->qp_stride, scr.data(), tb.data() + (size_t) q * hp.t_stride);
        }
    return 0;
}

int emul_expand(const cuipm_shape *sh, int cond_N, int nbatch, const double *qp, const double *sol2, double *sol, int nthreads)
{
    HostPlan hp;
    if (!build_plan(sh, cond_N, hp)) return -1;
    const Plan P = hp.plan(hp.ipool.data(), hp.upool.data());
    std::vector<double> scr(scratch_doubles(P));
    SeqExec ex{nthreads};
    for (int q = 0; q < nbatch; q++)
        expand_one(ex, P, qp + (size_t) q * hp.lo->qp_stride, sol2 + (size_t) q * hp.lc->sol_stride, sol + (size_t) q * hp.lo->sol_stride, scr.data());
    return 0;
}
}
