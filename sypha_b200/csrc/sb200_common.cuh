// sb200_common.cuh - shared device helpers, the device scalar block and the workspace layout.
//
// The scalar block replaces every host-side scalar of the reference's loop
// (/root/reference/src/sypha_solver.cpp:49-50: alpha, mu, muAff, sigma, alphaMaxPrim/Dual ...):
// kernels read and write it, the host only polls `done`.
#pragma once

#include <cuda_runtime.h>
#include <cfloat>
#include <cstdint>
#include <cstdio>
#include <string>
#include <vector>

#include "../../include/sypha_b200.h"

#define SB200_MAX_PARTIAL_BLOCKS 1184   // 148 SMs x 8
#define SB200_TRACE_ROWS 4096
#define SB200_TILE 64                   // dense M is padded to a multiple of this
#define SB200_NB 32                     // Cholesky panel width

namespace sb200 {

// ---------------------------------------------------------------------------------------------
// device scalar block
// ---------------------------------------------------------------------------------------------
struct Scalars
{
    double mu;
    double mu_aff;
    double sigma;
    double alpha_p;
    double alpha_d;
    double primal;
    double dual;
    double gap;
    double best_gap;
    unsigned long long amax_p;   // ordered-u64 encoding of min ratio (primal)
    unsigned long long amax_d;   // ordered-u64 encoding of min ratio (dual)
    unsigned long long min_x;    // starting point: min x~
    unsigned long long min_s;    // starting point: min s~
    double sum0, sum1, sum2;     // scratch sums (start point)
    // PCG
    double cg_rz, cg_pap, cg_rr, cg_rhs_norm2, cg_tol;
    int cg_iter, cg_done, cg_fail, cg_pad;
    long long cg_total;
    // control
    int iter;
    int stall;
    int done;
    int reason;
    int numerical;
    int chol_info;
    unsigned int ticket[8];
};

// parameters the kernels need every iteration (copied once per solve)
struct DevParams
{
    double eta;
    double mu_tol;
    double min_improv_ratio;
    int max_iter;
    int gap_enabled;
    int gap_window;
    int n_orig;
    int cg_max_iter;
    double cg_tol_initial;
    double cg_tol_final;
    double cg_tol_decay;
};

// ---------------------------------------------------------------------------------------------
// order-preserving double <-> u64 (so that atomicMin on u64 is an exact min on doubles)
// ---------------------------------------------------------------------------------------------
__host__ __device__ __forceinline__ unsigned long long ord_encode(double v)
{
#ifdef __CUDA_ARCH__
    unsigned long long b = (unsigned long long)__double_as_longlong(v);
#else
    unsigned long long b;
    memcpy(&b, &v, 8);
#endif
    return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}
__host__ __device__ __forceinline__ double ord_decode(unsigned long long k)
{
    unsigned long long b = (k >> 63) ? (k & 0x7fffffffffffffffull) : ~k;
#ifdef __CUDA_ARCH__
    return __longlong_as_double((long long)b);
#else
    double v;
    memcpy(&v, &b, 8);
    return v;
#endif
}
#define SB200_ORD_DBL_MAX 0xFFEFFFFFFFFFFFFFull   // ord_encode(DBL_MAX)

// ---------------------------------------------------------------------------------------------
// warp / block reductions (fixed shape => deterministic)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
        v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_min(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
        v = fmin(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
// sum over a thread block; result valid in thread 0.  `sh` needs 32 doubles.
__device__ __forceinline__ double block_sum(double v, double *sh)
{
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    v = warp_sum(v);
    __syncthreads();
    if (lane == 0)
        sh[w] = v;
    __syncthreads();
    if (w == 0)
    {
        const int nw = (blockDim.x + 31) >> 5;
        v = (lane < nw) ? sh[lane] : 0.0;
        v = warp_sum(v);
    }
    return v;
}
__device__ __forceinline__ double block_min(double v, double *sh)
{
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    v = warp_min(v);
    __syncthreads();
    if (lane == 0)
        sh[w] = v;
    __syncthreads();
    if (w == 0)
    {
        const int nw = (blockDim.x + 31) >> 5;
        v = (lane < nw) ? sh[lane] : DBL_MAX;
        v = warp_min(v);
    }
    return v;
}

// "last block" ticket: returns true in every thread of the block that arrives last.
__device__ __forceinline__ bool last_block_arrives(unsigned int *ticket, unsigned int nblocks)
{
    __shared__ bool is_last;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0)
    {
        unsigned int t = atomicAdd(ticket, 1u);
        is_last = (t == nblocks - 1);
        if (is_last)
            *ticket = 0u;    // re-arm for the next use (no other block touches it any more)
    }
    __syncthreads();
    if (is_last)
        __threadfence();
    return is_last;
}

// sum partial[0..nb) in a fixed order with one block; result valid in thread 0.
__device__ __forceinline__ double reduce_partials(const volatile double *partial, int nb, double *sh)
{
    double v = 0.0;
    for (int i = threadIdx.x; i < nb; i += blockDim.x)
        v += partial[i];
    return block_sum(v, sh);
}

// ---------------------------------------------------------------------------------------------
// host-side error handling: never exit(), record and return
// ---------------------------------------------------------------------------------------------
struct ErrorSink
{
    std::string msg;
};

#define SB200_CUDA_TRY(sink, call)                                                              \
    do                                                                                          \
    {                                                                                           \
        cudaError_t e__ = (call);                                                               \
        if (e__ != cudaSuccess)                                                                 \
        {                                                                                       \
            char buf__[512];                                                                    \
            snprintf(buf__, sizeof buf__, "%s:%d: %s -> %s", __FILE__, __LINE__, #call,         \
                     cudaGetErrorString(e__));                                                  \
            (sink).msg = buf__;                                                                 \
            return (e__ == cudaErrorMemoryAllocation) ? SB200_ERR_NOMEM : SB200_ERR_CUDA;       \
        }                                                                                       \
    } while (0)

inline int grid_for(long long n, int block, int cap = SB200_MAX_PARTIAL_BLOCKS)
{
    long long g = (n + block - 1) / block;
    if (g < 1) g = 1;
    if (g > cap) g = cap;
    return (int)g;
}

// ---------------------------------------------------------------------------------------------
// Programmatic dependent launch: the ~12 kernels of an IPM iteration are each tens of microseconds or less, and
// a plain stream (or graph) edge costs the launch latency of every one of them after its predecessor has drained.
// Launched with the programmatic-serialization attribute a kernel's CTAs are dispatched while the predecessor is
// still running; `pdl_wait()` - the FIRST statement of every kernel launched that way - blocks until the
// predecessor grid has completed and its writes are visible, so only the dispatch overlaps, never the data.
// `pdl_trigger()` right behind it lets the successor be dispatched in turn.  Kernels without these calls, and
// launches without the attribute, keep plain stream order.
// ---------------------------------------------------------------------------------------------
// MEASURED (B200, scpnrh shape, iteration replayed from a CUDA graph with programmatic edges - 13 kernel nodes):
// stand-alone launch groups get shorter (one solve 14.4 -> 12.2 us, fused vector kernels 14.9 -> 11.4 us) but the
// whole loop gets LONGER: 14.4 -> 15.2 ms per LP (2958 -> 2803 iter/s), and 15.4 ms when the factorisation also
// triggers early (successor CTAs parked on the SMs slow its pivot chains); B&B 1671 -> 1647 / 1347 nodes/s.
// Off by default; the calls below compile to nothing and launch_pdl is a plain launch.
#ifndef SB200_V_PDL
#define SB200_V_PDL 0
#endif
__device__ __forceinline__ void pdl_wait(bool trigger = true)
{
#if SB200_V_PDL
    asm volatile("griddepcontrol.wait;" ::: "memory");
    if (trigger) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
#else
    (void)trigger;
#endif
}
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args... args)
{
#if SB200_V_PDL
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
#else
    kernel<<<grid, block, smem, st>>>(KArgs(args)...);
    return cudaGetLastError();
#endif
}

// eight 2-byte column ids of one 16-byte chunk of the compact normal-matrix pattern: sum of d over them, d in shared
// memory (fixed association: the assembly kernel and the factorisation's in-task assembly give the same bits)
__device__ __forceinline__ double chunk_gather8_s(uint4 v, const double *ds)
{
    return ((ds[v.x & 0xffffu] + ds[v.x >> 16]) + (ds[v.y & 0xffffu] + ds[v.y >> 16])) +
           ((ds[v.z & 0xffffu] + ds[v.z >> 16]) + (ds[v.w & 0xffffu] + ds[v.w >> 16]));
}

// Function attributes (opt-in shared memory) are per DEVICE: a process that holds workspaces on several GPUs must
// set them once on each.  `seen` is a caller-owned bit mask (devices 0..63); returns true on the first call for
// the current device.
inline bool first_use_on_device(unsigned long long &seen)
{
    int dev = 0;
    cudaGetDevice(&dev);
    const unsigned long long bit = 1ull << (dev & 63);
    if (seen & bit) return false;
    seen |= bit;
    return true;
}

extern long long g_launch_count;   // kernels launched by this library (bench.py's gpu_launches)

} // namespace sb200
