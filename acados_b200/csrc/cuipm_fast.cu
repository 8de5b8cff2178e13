// cuipm_fast.cu -- CUDA instantiation (sm_100a) of the throughput kernel of the batched OCP-QP interior-point solver.
//
// The kernel body is cuipm_fast_core.h (a group of G lanes per QP, 32/G QPs per warp in lock step, register-tiled
// rank-k updates, stage blocks staged with asynchronous copies); this file binds its warp primitives to the hardware
// (shuffles, votes, cp.async, __syncwarp) and launches one warp per CTA, 32/G QPs per CTA.  Replaces, for the shapes
// listed in fast_available(), the reference's d_ocp_qp_ipm_solve (external/hpipm/ocp_qp/x_ocp_qp_ipm.c:2684-3120) on
// BLASFEO's panel-major kernels (dsyrk_dpotrf_ln_mn, dtrmm_rlnn: external/blasfeo/blasfeo_hp_pm/d_lapack_lib4.c:1513,
// d_blas3_lib4.c:4893).
#include <cuda_runtime.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>

#include "cuipm_device.h"

#define FK_DEV __device__ __forceinline__
static __device__ __forceinline__ int fk_lane() { return (int) (threadIdx.x & 31u); }
static __device__ __forceinline__ void fk_sync() { __syncwarp(); }
static __device__ __forceinline__ double fk_shfl_xor(double v, int m) { return __shfl_xor_sync(0xffffffffu, v, m); }
static __device__ __forceinline__ int fk_shfl_xor_i(int v, int m) { return __shfl_xor_sync(0xffffffffu, v, m); }
static __device__ __forceinline__ bool fk_any(bool p) { return __any_sync(0xffffffffu, p) != 0; }
// asynchronous global -> shared copies (LDGSTS): no register staging; 16-byte copies bypass L1 (each stage block is read
// once per sweep), completion through the per-thread group wait + a warp barrier
static __device__ __forceinline__ void fk_cp16(double *sdst, const double *gsrc)
{
    const unsigned sa = (unsigned) __cvta_generic_to_shared(sdst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sa), "l"(gsrc) : "memory");
}
static __device__ __forceinline__ void fk_cp8(double *sdst, const double *gsrc)
{
    const unsigned sa = (unsigned) __cvta_generic_to_shared(sdst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(sa), "l"(gsrc) : "memory");
}
static __device__ __forceinline__ void fk_cp_wait() { asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory"); }
static __device__ __forceinline__ double fk_ldg(const double *p) { return __ldg(p); }
static __device__ __forceinline__ double fk_rsqrt(double x) { return rsqrt(x); }
static __device__ __forceinline__ int fk_atomic_inc(int *p) { return atomicAdd(p, 1); }

#include "cuipm_fast_core.h"

namespace cuipm {

namespace {

extern __shared__ __align__(16) double g_fsmem[];

template <int NX, int NU, int G, int MINB>
__global__ void __launch_bounds__(32, MINB) cuipm_fast_kernel(const __grid_constant__ FastArgs A)
{
    using K = fastk::Ker<NX, NU, G>;
    K k(A, g_fsmem);
    int q = (int) blockIdx.x * K::QPW + k.gq;
    const bool valid = q < A.nbatch;
    if (!valid) q = A.nbatch - 1;
    k.solve(q, valid);
}

template <int NX, int NU, int G>
void sizes(FastArgs &F, int *qpw)
{
    using K = fastk::Ker<NX, NU, G>;
    F.vsize = fastk::vector_pool_doubles(NX, NX + NU, F.nce, F.nbe, F.ns2e, F.nve);
    int gs = K::MATS + F.vsize;
    // 32 / G groups share a warp: a group stride of 4 (mod 16) doubles spreads their broadcast loads over the banks
    while (gs % 16 != 4) gs++;
    F.gstride = gs;
    *qpw = K::QPW;
}

template <int NX, int NU, int G, int MINB>
cudaError_t launch_one(const FastArgs &F, cudaStream_t stream)
{
    using K = fastk::Ker<NX, NU, G>;
    const size_t smem = sizeof(double) * (size_t) F.gstride * K::QPW;
    cudaError_t err = cudaFuncSetAttribute(cuipm_fast_kernel<NX, NU, G, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem);
    if (err != cudaSuccess) return err;
    // the CTAs of one SM together need most of its shared memory: ask for the largest carve-out (the default heuristic
    // sized it for a single CTA, which left one warp per SM resident)
    err = cudaFuncSetAttribute(cuipm_fast_kernel<NX, NU, G, MINB>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    if (err != cudaSuccess) return err;
    if (getenv("CUIPM_DEBUG"))
    {
        int nblk = 0;
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nblk, cuipm_fast_kernel<NX, NU, G, MINB>, 32, smem);
        fprintf(stderr, "cuipm_fast_kernel<%d,%d,%d>: %zu bytes of shared memory per CTA, %d CTAs (%d QPs) per SM\n", NX, NU, G, smem, nblk, nblk * K::QPW);
    }
    const int grid = (F.nbatch + K::QPW - 1) / K::QPW;
    cuipm_fast_kernel<NX, NU, G, MINB><<<grid, 32, smem, stream>>>(F);
    return cudaGetLastError();
}

}  // namespace

// (interior nx, nu) pairs with an instance of the throughput kernel, and the lanes per QP each one runs with
#define CUIPM_FAST_INSTANCES(X) \
    X(21, 3, 8, 4)              \
    X(8, 3, 4, 8)               \
    X(4, 1, 2, 8)               \
    X(12, 4, 8, 6)

bool fast_available(int nx, int nu, FastArgs &F, int *qp_per_warp)
{
#define X(NX_, NU_, G_, MB_) \
    if (nx == NX_ && nu == NU_) { sizes<NX_, NU_, G_>(F, qp_per_warp); return sizeof(double) * (size_t) F.gstride * (32 / G_) <= 227 * 1024; }
    CUIPM_FAST_INSTANCES(X)
#undef X
    return false;
}

int launch_fast(const FastArgs &F, void *stream_)
{
    cudaStream_t stream = (cudaStream_t) stream_;
    const int nx = F.s1.nx, nu = F.s1.nu;
#define X(NX_, NU_, G_, MB_) \
    if (nx == NX_ && nu == NU_) return (int) launch_one<NX_, NU_, G_, MB_>(F, stream);
    CUIPM_FAST_INSTANCES(X)
#undef X
    return (int) cudaErrorInvalidValue;
}

}  // namespace cuipm
