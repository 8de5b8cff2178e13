// cuipm_condense.cu -- batched partial (block) condensing and expansion on the device: the CUDA instantiation of
// cuipm_condense_core.h (one CTA per QP, a phase = all threads of the CTA + __syncthreads) and its C-ABI entry points.
// Reference: ocp_qp_partial_condensing (acados/ocp_qp/ocp_qp_partial_condensing.c:523-689) -> d_part_cond_qp_cond /
// d_part_cond_qp_expand_sol (external/hpipm/cond/x_part_cond.c:410-866).
#include <cuda_runtime.h>

#include <string>

#include "cuipm.h"
#include "cuipm_condense_plan.h"
#include "cuipm_internal.h"

using namespace cuipm;
using namespace cuipm_cond;

namespace {

struct CtaExec
{
    __device__ int nthreads() const { return (int) blockDim.x; }
    template <class F>
    __device__ void phase(F f)
    {
        f((int) threadIdx.x);
        __syncthreads();
    }
};

template <int MODE>
__global__ void condense_kernel(Plan P, const double *qp, double *qp2, double *tbuf, int nbatch)
{
    extern __shared__ double scr[];
    const int q = blockIdx.x;
    if (q >= nbatch) return;
    CtaExec ex;
    condense_one<MODE>(ex, P, qp + (size_t) q * P.o.qp_stride, qp2 + (size_t) q * P.c.qp_stride, scr,
                       MODE == COND_ALL ? nullptr : tbuf + (size_t) q * P.t_stride);
}

__global__ void expand_kernel(Plan P, const double *qp, const double *sol2, double *sol, int nbatch)
{
    extern __shared__ double scr[];
    const int q = blockIdx.x;
    if (q >= nbatch) return;
    CtaExec ex;
    expand_one(ex, P, qp + (size_t) q * P.o.qp_stride, sol2 + (size_t) q * P.c.sol_stride, sol + (size_t) q * P.o.sol_stride, scr);
}

}  // namespace

struct cuipm_condenser
{
    int device = 0;
    HostPlan hp;
    Plan P{};
    int *d_i = nullptr;
    unsigned *d_u = nullptr;
    size_t smem = 0;
    double *d_t = nullptr;        // T_j of the QPs of the last lhs pass (t_stride doubles per QP)
    int t_cap = 0, t_valid = 0;   // QPs the buffer holds / QPs the last lhs pass filled
};

#define CKC(call)                                                                                       \
    do {                                                                                                \
        cudaError_t e_ = (call);                                                                        \
        if (e_ != cudaSuccess)                                                                          \
        {                                                                                               \
            set_error(std::string(#call) + ": " + cudaGetErrorString(e_));                              \
            return CUIPM_ERR_CUDA;                                                                      \
        }                                                                                               \
    } while (0)

extern "C" void cuipm_condenser_destroy(cuipm_condenser *c)
{
    if (!c) return;
    cudaSetDevice(c->device);
    cudaFree(c->d_i);
    cudaFree(c->d_u);
    cudaFree(c->d_t);
    delete c;
}

extern "C" cuipm_condenser *cuipm_condenser_create(const cuipm_shape *shape, int cond_N, int device)
{
    if (!shape) { set_error("cuipm_condenser_create: bad arguments"); return nullptr; }
    cuipm_condenser *c = new cuipm_condenser();
    c->device = device;
    if (!build_plan(shape, cond_N, c->hp))
    {
        set_error("cuipm_condenser_create: need 1 <= cond_N <= N (and records within 32-bit offsets)");
        delete c;
        return nullptr;
    }
    if (cudaSetDevice(device) != cudaSuccess || cudaMalloc(&c->d_i, sizeof(int) * c->hp.ipool.size()) != cudaSuccess
        || cudaMalloc(&c->d_u, sizeof(unsigned) * c->hp.upool.size()) != cudaSuccess
        || cudaMemcpy(c->d_i, c->hp.ipool.data(), sizeof(int) * c->hp.ipool.size(), cudaMemcpyHostToDevice) != cudaSuccess
        || cudaMemcpy(c->d_u, c->hp.upool.data(), sizeof(unsigned) * c->hp.upool.size(), cudaMemcpyHostToDevice) != cudaSuccess)
    {
        set_error("cuipm_condenser_create: CUDA allocation failed (no CPU fallback)");
        cuipm_condenser_destroy(c);
        return nullptr;
    }
    c->P = c->hp.plan(c->d_i, c->d_u);
    c->smem = sizeof(double) * (size_t) scratch_doubles(c->P);
    if (c->smem > 227 * 1024) { set_error("condensed stage too large for the shared-memory scratch"); cuipm_condenser_destroy(c); return nullptr; }
    cudaFuncSetAttribute(condense_kernel<COND_ALL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) c->smem);
    cudaFuncSetAttribute(condense_kernel<COND_LHS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) c->smem);
    cudaFuncSetAttribute(condense_kernel<COND_RHS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) c->smem);
    cudaFuncSetAttribute(expand_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) c->smem);
    return c;
}

extern "C" const cuipm_shape *cuipm_condenser_condensed_shape(const cuipm_condenser *c) { return c ? &c->hp.cshape : nullptr; }

extern "C" int cuipm_condense_device(cuipm_condenser *c, int nbatch, const double *d_qp, double *d_qp_cond, void *stream)
{
    if (!c || nbatch < 0 || !d_qp || !d_qp_cond) { set_error("cuipm_condense_device: bad arguments"); return CUIPM_ERR_INVALID; }
    if (nbatch == 0) return CUIPM_OK;
    CKC(cudaSetDevice(c->device));
    // the masks of a fresh record are 1 and untouched entries of d / Z / z are 0 in the reference's layout: clear, then fill
    CKC(cudaMemsetAsync(d_qp_cond, 0, sizeof(double) * c->hp.lc->qp_stride * (size_t) nbatch, (cudaStream_t) stream));
    condense_kernel<COND_ALL><<<nbatch, 128, c->smem, (cudaStream_t) stream>>>(c->P, d_qp, d_qp_cond, nullptr, nbatch);
    CKC(cudaGetLastError());
    return CUIPM_OK;
}

// The split of acados' xcond solver (ocp_qp_xcond_solver.c:591-669: condense_lhs in the preparation phase of an SQP-RTI step,
// condense_rhs_and_solve in its feedback phase): the lhs pass condenses the whole QP and keeps the prediction matrices T_j of
// every stage per QP on the device; the rhs pass recomputes the vectors of the condensed records (gradient, dynamics offset,
// shifted bounds, slack gradient) for new b, rq, d, z of the same matrices -- O(nx^2 + nx n2) per stage instead of O(nx^2 n2).
extern "C" int cuipm_condense_lhs_device(cuipm_condenser *c, int nbatch, const double *d_qp, double *d_qp_cond, void *stream)
{
    if (!c || nbatch < 0 || !d_qp || !d_qp_cond) { set_error("cuipm_condense_lhs_device: bad arguments"); return CUIPM_ERR_INVALID; }
    if (nbatch == 0) return CUIPM_OK;
    CKC(cudaSetDevice(c->device));
    if (c->t_cap < nbatch)
    {
        CKC(cudaStreamSynchronize((cudaStream_t) stream));
        cudaFree(c->d_t);
        c->d_t = nullptr; c->t_cap = 0; c->t_valid = 0;
        CKC(cudaMalloc(&c->d_t, sizeof(double) * (size_t) c->P.t_stride * (size_t) nbatch));
        c->t_cap = nbatch;
    }
    CKC(cudaMemsetAsync(d_qp_cond, 0, sizeof(double) * c->hp.lc->qp_stride * (size_t) nbatch, (cudaStream_t) stream));
    condense_kernel<COND_LHS><<<nbatch, 128, c->smem, (cudaStream_t) stream>>>(c->P, d_qp, d_qp_cond, c->d_t, nbatch);
    CKC(cudaGetLastError());
    c->t_valid = nbatch;
    return CUIPM_OK;
}

extern "C" int cuipm_condense_rhs_device(cuipm_condenser *c, int nbatch, const double *d_qp, double *d_qp_cond, void *stream)
{
    if (!c || nbatch < 0 || !d_qp || !d_qp_cond) { set_error("cuipm_condense_rhs_device: bad arguments"); return CUIPM_ERR_INVALID; }
    if (nbatch > c->t_valid) { set_error("cuipm_condense_rhs_device: no lhs pass for this many QPs (call cuipm_condense_lhs_device first)"); return CUIPM_ERR_INVALID; }
    if (nbatch == 0) return CUIPM_OK;
    CKC(cudaSetDevice(c->device));
    condense_kernel<COND_RHS><<<nbatch, 128, c->smem, (cudaStream_t) stream>>>(c->P, d_qp, d_qp_cond, c->d_t, nbatch);
    CKC(cudaGetLastError());
    return CUIPM_OK;
}

extern "C" int cuipm_expand_device(cuipm_condenser *c, int nbatch, const double *d_qp, const double *d_sol_cond, double *d_sol, void *stream)
{
    if (!c || nbatch < 0 || !d_qp || !d_sol_cond || !d_sol) { set_error("cuipm_expand_device: bad arguments"); return CUIPM_ERR_INVALID; }
    if (nbatch == 0) return CUIPM_OK;
    CKC(cudaSetDevice(c->device));
    CKC(cudaMemsetAsync(d_sol, 0, sizeof(double) * c->hp.lo->sol_stride * (size_t) nbatch, (cudaStream_t) stream));
    expand_kernel<<<nbatch, 128, c->smem, (cudaStream_t) stream>>>(c->P, d_qp, d_sol_cond, d_sol, nbatch);
    CKC(cudaGetLastError());
    return CUIPM_OK;
}
