// sb200_pcg.cu - matrix-free Jacobi-preconditioned CG on the normal equations (A D A') dy = rhs.
//
// Replaces /root/reference/src/sypha_solver_krylov.cu:228-392 (krylovSolveCG): per CG iteration the
// reference issues 2 cuSPARSE SpMVs (one on the transposed CSR), a scale kernel, 3 blocking
// host-pointer reductions and 5 BLAS-1 launches.  Here one CG iteration is 4 launches and no host
// round trip: q = D A'p (CSC gather, sb200_spmv.cu) -> Ap = A q fused with p.Ap -> fused
// x/r/z update with ||r||^2 and r.z -> p update.  alpha, beta, the convergence test and the failure
// tests live in the device scalar block; the host polls `cg_done` once per chunk of iterations.
// Same recurrences, tolerance schedule and failure rules as the reference (x0 = 0, relative residual
// ||r||/||rhs|| < tol, failure on pAp <= 0 / non-finite / |rz| < 1e-30 / iteration cap).
#include "sb200_kernels.cuh"
#include "sb200_pcg.cuh"

namespace sb200 {

static constexpr int kBlock = 256;

// r = rhs ; x = 0 ; z = r/diag ; p = z ; rz = r.z ; ||rhs||^2 ; tolerance for this solve
__global__ void k_cg_init(PcgVecs C, Scalars *sc, const DevParams *__restrict__ Pp, double fixed_tol,
                          int honour_done)
{
    __shared__ double sh[32];
    if (honour_done && sc->done) return;
    double a_rz = 0.0, a_nn = 0.0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < C.m; i += gridDim.x * blockDim.x)
    {
        const double r = C.rhs[i];
        const double z = r / fmax(C.diag[i], 1e-30);
        C.x[i] = 0.0;
        C.r[i] = r;
        C.z[i] = z;
        C.p[i] = z;
        a_rz += r * z;
        a_nn += r * r;
    }
    a_rz = block_sum(a_rz, sh);
    a_nn = block_sum(a_nn, sh);
    if (threadIdx.x == 0)
    {
        C.partial[blockIdx.x] = a_rz;
        C.partial[SB200_MAX_PARTIAL_BLOCKS + blockIdx.x] = a_nn;
    }
    if (last_block_arrives(&sc->ticket[4], gridDim.x))
    {
        const double t_rz = reduce_partials(C.partial, gridDim.x, sh);
        const double t_nn = reduce_partials(C.partial + SB200_MAX_PARTIAL_BLOCKS, gridDim.x, sh);
        if (threadIdx.x == 0)
        {
            const DevParams P = *Pp;
            sc->cg_rz = t_rz;
            sc->cg_rhs_norm2 = t_nn;
            sc->cg_iter = 0;
            sc->cg_fail = 0;
            // krylov.cu:243-250: zero right-hand side => zero solution, 0 iterations
            sc->cg_done = (sqrt(t_nn) < 1e-30) ? 1 : 0;
            sc->cg_tol = (fixed_tol > 0.0)
                             ? fixed_tol
                             : fmax(P.cg_tol_final, P.cg_tol_initial * pow(P.cg_tol_decay, (double)sc->iter));
        }
    }
}

// Ap = A q, fused with the partial sums of p.Ap; last block publishes pAp and the failure test
__global__ void __launch_bounds__(256)
k_cg_matvec(CsrView A, PcgVecs C, Scalars *sc)
{
    __shared__ double sh[32];
    if (sc->cg_done) return;
    const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
    double dot = 0.0;
    for (int row = blockIdx.x * wpb + (threadIdx.x >> 5); row < A.m; row += gridDim.x * wpb)
    {
        const int a = A.offs[row], e = A.offs[row + 1];
        double acc = 0.0;
        for (int k = a + lane; k < e; k += 32)
            acc += A.vals[k] * __ldg(C.q + A.inds[k]);
        acc = warp_sum(acc);
        if (lane == 0)
        {
            C.Ap[row] = acc;
            dot += C.p[row] * acc;
        }
    }
    dot = block_sum(dot, sh);
    if (threadIdx.x == 0) C.partial[blockIdx.x] = dot;
    if (last_block_arrives(&sc->ticket[5], gridDim.x))
    {
        const double pap = reduce_partials(C.partial, gridDim.x, sh);
        if (threadIdx.x == 0)
        {
            sc->cg_pap = pap;
            if (!(pap > 0.0) || !isfinite(pap))     // krylov.cu:335-339
            {
                sc->cg_fail = 1;
                sc->cg_done = 1;
            }
        }
    }
}

// x += a p ; r -= a Ap ; z = r/diag ; ||r||^2, r.z ; convergence / failure / beta
__global__ void k_cg_update(PcgVecs C, Scalars *sc, const DevParams *__restrict__ Pp, int max_iter_override)
{
    __shared__ double sh[32];
    if (sc->cg_done) return;
    const double alpha = sc->cg_rz / sc->cg_pap;
    double a_rr = 0.0, a_rz = 0.0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < C.m; i += gridDim.x * blockDim.x)
    {
        C.x[i] += alpha * C.p[i];
        const double r = C.r[i] - alpha * C.Ap[i];
        const double z = r / fmax(C.diag[i], 1e-30);
        C.r[i] = r;
        C.z[i] = z;
        a_rr += r * r;
        a_rz += r * z;
    }
    a_rr = block_sum(a_rr, sh);
    a_rz = block_sum(a_rz, sh);
    if (threadIdx.x == 0)
    {
        C.partial[blockIdx.x] = a_rr;
        C.partial[SB200_MAX_PARTIAL_BLOCKS + blockIdx.x] = a_rz;
    }
    if (last_block_arrives(&sc->ticket[6], gridDim.x))
    {
        const double t_rr = reduce_partials(C.partial, gridDim.x, sh);
        const double t_rz = reduce_partials(C.partial + SB200_MAX_PARTIAL_BLOCKS, gridDim.x, sh);
        if (threadIdx.x == 0)
        {
            const int cap = max_iter_override > 0 ? max_iter_override : Pp->cg_max_iter;
            const int it = sc->cg_iter + 1;
            sc->cg_iter = it;
            sc->cg_total += 1;
            sc->cg_rr = t_rr;
            int done = 0;
            if (sqrt(t_rr) / sqrt(sc->cg_rhs_norm2) < sc->cg_tol)
                done = 1;                                  // krylov.cu:352-357
            else if (fabs(sc->cg_rz) < 1e-30)
            {
                sc->cg_fail = 1;                           // krylov.cu:369-373
                done = 1;
            }
            else if (it >= cap)
            {
                sc->cg_fail = 1;                           // krylov.cu:389-390
                done = 1;
            }
            else
            {
                sc->sum0 = t_rz / sc->cg_rz;               // beta
                sc->cg_rz = t_rz;
            }
            __threadfence();
            sc->cg_done = done;
        }
    }
}

// p = z + beta p
__global__ void k_cg_direction(PcgVecs C, const Scalars *sc)
{
    if (sc->cg_done) return;
    const double beta = sc->sum0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < C.m; i += gridDim.x * blockDim.x)
        C.p[i] = C.z[i] + beta * C.p[i];
}

// a CG failure inside the IPM loop is an LP failure (sypha_solver.cpp:558-566)
__global__ void k_cg_check(Scalars *sc)
{
    if (sc->cg_fail && !sc->done)
    {
        sc->numerical = 1;
        sc->reason = SB200_TERM_INFEASIBLE_OR_NUMERICAL;
        sc->done = 1;
    }
}

void launch_cg_init(const PcgVecs &C, Scalars *sc, const DevParams *P, double fixed_tol, int honour_done,
                    cudaStream_t st)
{
    k_cg_init<<<grid_for(C.m, kBlock), kBlock, 0, st>>>(C, sc, P, fixed_tol, honour_done);
    ++g_launch_count;
}

void launch_cg_iteration(const CsrView &A, const CscView &At, const PcgVecs &C, const IpmVecs &V,
                         const DevParams *P, int max_iter_override, cudaStream_t st)
{
    // q = D A'p : CSC gather with the scale epilogue; V.d holds D; early exit on cg_done
    launch_spmv_csc_cg(At, C.p, C.q, C.dscale, V.sc, st);
    if (A.blk)
        launch_blk_cg_matvec(*A.blk, C, V.sc, st);
    else
    {
        k_cg_matvec<<<grid_for((long long)A.m * 32, 256), 256, 0, st>>>(A, C, V.sc);
        ++g_launch_count;
    }
    k_cg_update<<<grid_for(C.m, kBlock), kBlock, 0, st>>>(C, V.sc, P, max_iter_override);
    k_cg_direction<<<grid_for(C.m, kBlock), kBlock, 0, st>>>(C, V.sc);
    g_launch_count += 2;
}

void launch_cg_check(Scalars *sc, cudaStream_t st)
{
    k_cg_check<<<1, 1, 0, st>>>(sc);
    ++g_launch_count;
}

} // namespace sb200
