// sb200_vector.cu - fused elementwise / reduction kernels of the Mehrotra loop.
//
// Replaces /root/reference/src/sypha_solver_utils.cu (5 kernels + 2 n-sized temporaries + a D2H per
// ratio test) and the ~20 cuBLAS level-1 calls per iteration in
// /root/reference/src/sypha_solver.cpp:505,596-629,693-753.  Every scalar lives in the device
// `Scalars` block; reductions are "partials + last block" with a fixed summation order, so the
// results are deterministic run to run.  All kernels are HBM-bound streaming kernels: grid-stride,
// 256 threads, grid capped at 148*8 blocks.
#include "sb200_kernels.cuh"

namespace sb200 {

long long g_launch_count = 0;

static constexpr int kBlock = 256;

// ---------------------------------------------------------------------------------------------
// L0 drop-ins (reference: sypha_solver_utils.cu:5-17, :51-65, :68-177)
// ---------------------------------------------------------------------------------------------
__global__ void k_elem_min_mult(const double *__restrict__ x, const double *__restrict__ s,
                                double *__restrict__ out, int n)
{
    for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < n; j += gridDim.x * blockDim.x)
        out[j] = -x[j] * s[j];
}

__global__ void k_corrector_rhs(const double *__restrict__ dx, const double *__restrict__ ds,
                                double sigma, double mu, double *__restrict__ out, int n)
{
    const double sm = sigma * mu;
    for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < n; j += gridDim.x * blockDim.x)
        out[j] = -dx[j] * ds[j] + sm;
}

// single pass: ratio + block min + one atomicMin per block on the ordered-u64 encoding
__global__ void k_alpha_max(const double *__restrict__ x, const double *__restrict__ dx,
                            const double *__restrict__ s, const double *__restrict__ ds, int n,
                            unsigned long long *__restrict__ ord2)
{
    __shared__ double sh[32];
    double mp = DBL_MAX, md = DBL_MAX;
    for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < n; j += gridDim.x * blockDim.x)
    {
        const double dxj = dx[j], dsj = ds[j];
        if (dxj < 0.0) mp = fmin(mp, -x[j] / dxj);      // strict <, utils.cu:77-78
        if (dsj < 0.0) md = fmin(md, -s[j] / dsj);
    }
    mp = block_min(mp, sh);
    md = block_min(md, sh);
    if (threadIdx.x == 0)
    {
        atomicMin(&ord2[0], ord_encode(mp));
        atomicMin(&ord2[1], ord_encode(md));
    }
}
__global__ void k_ord_init2(unsigned long long *ord2)
{
    ord2[0] = SB200_ORD_DBL_MAX;
    ord2[1] = SB200_ORD_DBL_MAX;
}
__global__ void k_ord_decode2(const unsigned long long *ord2, double *out)
{
    out[0] = ord_decode(ord2[0]);
    out[1] = ord_decode(ord2[1]);
}

void launch_elem_min_mult(const double *x, const double *s, double *out, int n, cudaStream_t st)
{
    k_elem_min_mult<<<grid_for(n, kBlock), kBlock, 0, st>>>(x, s, out, n);
    ++g_launch_count;
}
void launch_corrector_rhs(const double *dx, const double *ds, double sigma, double mu, double *out,
                          int n, cudaStream_t st)
{
    k_corrector_rhs<<<grid_for(n, kBlock), kBlock, 0, st>>>(dx, ds, sigma, mu, out, n);
    ++g_launch_count;
}
void launch_alpha_max(const double *x, const double *dx, const double *s, const double *ds, int n,
                      unsigned long long *d_ord2, double *d_result, cudaStream_t st)
{
    k_ord_init2<<<1, 1, 0, st>>>(d_ord2);
    k_alpha_max<<<grid_for(n, kBlock), kBlock, 0, st>>>(x, dx, s, ds, n, d_ord2);
    k_ord_decode2<<<1, 1, 0, st>>>(d_ord2, d_result);
    g_launch_count += 3;
}

__global__ void k_fill(double *p, double v, long long n)
{
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
         i += (long long)gridDim.x * blockDim.x)
        p[i] = v;
}
void launch_fill(double *p, double v, long long n, cudaStream_t st)
{
    k_fill<<<grid_for(n, kBlock), kBlock, 0, st>>>(p, v, n);
    ++g_launch_count;
}

// ---------------------------------------------------------------------------------------------
// loop kernels over the scalar block
// ---------------------------------------------------------------------------------------------
__global__ void k_reset_scalars(Scalars *sc)
{
    Scalars z;
    memset(&z, 0, sizeof z);
    z.best_gap = INFINITY;
    z.amax_p = z.amax_d = z.min_x = z.min_s = SB200_ORD_DBL_MAX;
    *sc = z;
}
void launch_reset_scalars(Scalars *sc, cudaStream_t st)
{
    k_reset_scalars<<<1, 1, 0, st>>>(sc);
    ++g_launch_count;
}

// d = x/s ; resXS = -x.*s ; t = (x.*resC - resXS)/s      (sypha_solver.cpp:505, krylov.cu:19-24,55-63)
__device__ __forceinline__ void prologue_elem(const IpmVecs &V, int j, double xj, double sj, double rc)
{
    const double rxs = -xj * sj;
    V.resXS[j] = rxs;
    V.d[j] = xj / sj;
    V.t[j] = (xj * rc - rxs) / sj;
}

__global__ void k_prologue(IpmVecs V)
{
    if (V.sc->done) return;
    for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < V.n; j += gridDim.x * blockDim.x)
        prologue_elem(V, j, V.x[j], V.s[j], V.resC[j]);
}
void launch_prologue(const IpmVecs &V, cudaStream_t st)
{
    k_prologue<<<grid_for(V.n, kBlock), kBlock, 0, st>>>(V);
    ++g_launch_count;
}

// mu = x.s/n ; arm the loop      (sypha_solver.cpp:458-459, loop test :496)
__global__ void k_init_mu(IpmVecs V, const DevParams *__restrict__ Pp)
{
    __shared__ double sh[32];
    const DevParams P = *Pp;
    double acc = 0.0, a_p = 0.0, a_d = 0.0;
    const int lim = max(V.n, V.m);
    for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < lim; j += gridDim.x * blockDim.x)
    {
        if (j < V.n)
        {
            acc += V.x[j] * V.s[j];
            if (j < P.n_orig) a_p += V.x[j] * V.c[j];
        }
        if (j < V.m) a_d += V.y[j] * V.b[j];
    }
    acc = block_sum(acc, sh);
    a_p = block_sum(a_p, sh);
    a_d = block_sum(a_d, sh);
    if (threadIdx.x == 0)
    {
        V.partial[blockIdx.x] = acc;
        V.partial[SB200_MAX_PARTIAL_BLOCKS + blockIdx.x] = a_p;
        V.partial[2 * SB200_MAX_PARTIAL_BLOCKS + blockIdx.x] = a_d;
    }
    if (last_block_arrives(&V.sc->ticket[0], gridDim.x))
    {
        double tot = reduce_partials(V.partial, gridDim.x, sh);
        const double t_p = reduce_partials(V.partial + SB200_MAX_PARTIAL_BLOCKS, gridDim.x, sh);
        const double t_d = reduce_partials(V.partial + 2 * SB200_MAX_PARTIAL_BLOCKS, gridDim.x, sh);
        if (threadIdx.x == 0)
        {
            Scalars *sc = V.sc;
            const double mu = tot / (double)V.n;
            sc->mu = mu;
            sc->primal = t_p;
            sc->dual = t_d;
            sc->gap = fabs(t_p - t_d) / fmax(1.0, fabs(t_p));
            sc->iter = 0;
            sc->stall = 0;
            sc->best_gap = INFINITY;
            sc->numerical = 0;
            sc->reason = SB200_TERM_MAX_ITER;
            sc->amax_p = sc->amax_d = SB200_ORD_DBL_MAX;
            if (!(mu > P.mu_tol) || P.max_iter <= 0)
            {
                sc->done = 1;
                sc->reason = (mu <= P.mu_tol) ? SB200_TERM_CONVERGED : SB200_TERM_MAX_ITER;
            }
            else
                sc->done = 0;
        }
    }
}
void launch_init_mu(const IpmVecs &V, const DevParams *P, cudaStream_t st)
{
    k_init_mu<<<grid_for(max(V.n, V.m), kBlock), kBlock, 0, st>>>(V, P);
    ++g_launch_count;
}

// affine step lengths, mu_aff and sigma     (sypha_solver.cpp:596-622)
__global__ void k_affine_mu(IpmVecs V)
{
    pdl_wait();
    __shared__ double sh[32];
    Scalars *sc = V.sc;
    if (sc->done) return;
    const double ap = fmin(1.0, ord_decode(sc->amax_p));
    const double ad = fmin(1.0, ord_decode(sc->amax_d));
    double acc = 0.0;
    for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < V.n; j += gridDim.x * blockDim.x)
        acc += (V.x[j] + ap * V.dx[j]) * (V.s[j] + ad * V.ds[j]);
    acc = block_sum(acc, sh);
    if (threadIdx.x == 0) V.partial[blockIdx.x] = acc;
    if (last_block_arrives(&sc->ticket[1], gridDim.x))
    {
        double tot = reduce_partials(V.partial, gridDim.x, sh);
        if (threadIdx.x == 0)
        {
            const double mu_aff = tot / (double)V.n;
            const double r = mu_aff / sc->mu;
            sc->mu_aff = mu_aff;
            sc->sigma = r * r * r;                        // gsl_pow_3, :622
            sc->amax_p = sc->amax_d = SB200_ORD_DBL_MAX;  // re-arm the ratio test
        }
    }
}
void launch_affine_mu(const IpmVecs &V, cudaStream_t st)
{
    launch_pdl(k_affine_mu, grid_for(V.n, kBlock), kBlock, 0, st, V);
    ++g_launch_count;
}

// Tiny LPs (scp4x: n = 1200): the step-length / mu_aff / sigma reduction and the corrector right-hand side
// are ONE single-CTA kernel - no partials, no last-block ticket, one launch instead of two.  Not beyond a
// few thousand columns: one SM pulls ~126 GB/s from L2, and at n = 11000 the single-CTA pair measured
// 19.1 us against 6.4 + 4.3 us for the multi-CTA kernels.
#ifndef SB200_SINGLE_CTA_MAX
#define SB200_SINGLE_CTA_MAX 4096
#endif
static constexpr int kSingleCtaMax = SB200_SINGLE_CTA_MAX;
static constexpr int kSingleCtaThreads = 1024;
__global__ void __launch_bounds__(kSingleCtaThreads) k_affine_corrector(IpmVecs V)
{
    __shared__ double sh[32];
    __shared__ double s_sm;
    Scalars *sc = V.sc;
    if (sc->done) return;
    const double ap = fmin(1.0, ord_decode(sc->amax_p));
    const double ad = fmin(1.0, ord_decode(sc->amax_d));
    double acc = 0.0;
    for (int j = threadIdx.x; j < V.n; j += blockDim.x)
        acc += (V.x[j] + ap * V.dx[j]) * (V.s[j] + ad * V.ds[j]);
    acc = block_sum(acc, sh);
    if (threadIdx.x == 0)
    {
        const double mu_aff = acc / (double)V.n;
        const double r = mu_aff / sc->mu;
        sc->mu_aff = mu_aff;
        sc->sigma = r * r * r;                        // gsl_pow_3, :622
        sc->amax_p = sc->amax_d = SB200_ORD_DBL_MAX;  // re-arm the ratio test
        s_sm = r * r * r * sc->mu;
    }
    __syncthreads();
    const double sm = s_sm;
    for (int j = threadIdx.x; j < V.n; j += blockDim.x)
    {
        const double corr = -V.dx[j] * V.ds[j] + sm;
        const double rxs = V.resXS[j] + corr;
        V.resXS[j] = rxs;
        V.t[j] = (V.x[j] * V.resC[j] - rxs) / V.s[j];
    }
}

// resXS += -dxa.*dsa + sigma*mu ; t = (x.*resC - resXS)/s      (sypha_solver.cpp:625-629)
__global__ void k_corrector(IpmVecs V)
{
    pdl_wait();
    const Scalars *sc = V.sc;
    if (sc->done) return;
    const double sm = sc->sigma * sc->mu;
    for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < V.n; j += gridDim.x * blockDim.x)
    {
        const double corr = -V.dx[j] * V.ds[j] + sm;
        const double rxs = V.resXS[j] + corr;
        V.resXS[j] = rxs;
        V.t[j] = (V.x[j] * V.resC[j] - rxs) / V.s[j];
    }
}
void launch_corrector(const IpmVecs &V, cudaStream_t st)
{
    launch_pdl(k_corrector, grid_for(V.n, kBlock), kBlock, 0, st, V);
    ++g_launch_count;
}
void launch_affine_corrector(const IpmVecs &V, cudaStream_t st)
{
    if (V.n <= kSingleCtaMax)
    {
        k_affine_corrector<<<1, kSingleCtaThreads, 0, st>>>(V);
        ++g_launch_count;
        return;
    }
    launch_affine_mu(V, st);
    launch_corrector(V, st);
}

// step, residual scaling, mu / objectives / termination, and the next iteration's prologue
// (sypha_solver.cpp:693-769 and :505 of the following iteration)
__global__ void k_update(IpmVecs V, const DevParams *__restrict__ Pp)
{
    pdl_wait();
    __shared__ double sh[32];
    Scalars *sc = V.sc;
    if (sc->done) return;
    const DevParams P = *Pp;
    const double ap = fmin(1.0, P.eta * ord_decode(sc->amax_p));
    const double ad = fmin(1.0, P.eta * ord_decode(sc->amax_d));
    const double fc = -(ad - 1.0), fb = -(ap - 1.0);
    double a_xs = 0.0, a_p = 0.0, a_d = 0.0;
    const int lim = max(V.n, V.m);
    for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < lim; j += gridDim.x * blockDim.x)
    {
        if (j < V.n)
        {
            const double xj = V.x[j] + ap * V.dx[j];
            const double sj = V.s[j] + ad * V.ds[j];
            const double rc = V.resC[j] * fc;
            V.x[j] = xj;
            V.s[j] = sj;
            V.resC[j] = rc;
            a_xs += xj * sj;
            if (j < P.n_orig) a_p += xj * V.c[j];
            prologue_elem(V, j, xj, sj, rc);
        }
        if (j < V.m)
        {
            const double yj = V.y[j] + ad * V.dy[j];
            V.y[j] = yj;
            V.resB[j] *= fb;
            a_d += yj * V.b[j];
        }
    }
    a_xs = block_sum(a_xs, sh);
    a_p = block_sum(a_p, sh);
    a_d = block_sum(a_d, sh);
    if (threadIdx.x == 0)
    {
        V.partial[blockIdx.x] = a_xs;
        V.partial[SB200_MAX_PARTIAL_BLOCKS + blockIdx.x] = a_p;
        V.partial[2 * SB200_MAX_PARTIAL_BLOCKS + blockIdx.x] = a_d;
    }
    if (last_block_arrives(&sc->ticket[2], gridDim.x))
    {
        const double t_xs = reduce_partials(V.partial, gridDim.x, sh);
        const double t_p = reduce_partials(V.partial + SB200_MAX_PARTIAL_BLOCKS, gridDim.x, sh);
        const double t_d = reduce_partials(V.partial + 2 * SB200_MAX_PARTIAL_BLOCKS, gridDim.x, sh);
        if (threadIdx.x == 0)
        {
            const double mu_in = sc->mu;
            const double mu = t_xs / (double)V.n;
            const double gap = fabs(t_p - t_d) / fmax(1.0, fabs(t_p));
            const int it = sc->iter;
            if (it < SB200_TRACE_ROWS)
            {
                double *tr = V.trace + (size_t)it * SB200_TRACE_COLS;
                tr[0] = mu_in; tr[1] = mu; tr[2] = sc->mu_aff; tr[3] = sc->sigma;
                tr[4] = ap; tr[5] = ad; tr[6] = t_p; tr[7] = t_d;
            }
            sc->alpha_p = ap;
            sc->alpha_d = ad;
            sc->mu = mu;
            sc->primal = t_p;
            sc->dual = t_d;
            sc->gap = gap;
            sc->amax_p = sc->amax_d = SB200_ORD_DBL_MAX;
            int done = 0;
            if (!isfinite(mu) || mu < 0.0 || !isfinite(t_p) || !isfinite(t_d) || !isfinite(gap))
            {   // :723-729, :748-753 (iteration not counted)
                sc->numerical = 1;
                sc->reason = SB200_TERM_INFEASIBLE_OR_NUMERICAL;
                done = 1;
            }
            else
            {
                if (gap < sc->best_gap * (1.0 - P.min_improv_ratio))
                {
                    sc->best_gap = gap;
                    sc->stall = 0;
                }
                else if (P.gap_enabled)
                {
                    if (++sc->stall >= P.gap_window)
                    {
                        sc->reason = SB200_TERM_GAP_STALLED;
                        done = 1;
                    }
                }
                sc->iter = it + 1;
                if (!done && (it + 1 >= P.max_iter || !(mu > P.mu_tol)))
                {
                    sc->reason = (mu <= P.mu_tol) ? SB200_TERM_CONVERGED : SB200_TERM_MAX_ITER;
                    done = 1;
                }
            }
            __threadfence();
            sc->done = done;
        }
    }
}
void launch_update(const IpmVecs &V, const DevParams *P, cudaStream_t st)
{
    const int lim = max(V.n, V.m);
    if (lim <= kSingleCtaMax)       // one CTA: the "last block" is the only block
        k_update<<<1, kSingleCtaThreads, 0, st>>>(V, P);
    else
        launch_pdl(k_update, grid_for(lim, kBlock), kBlock, 0, st, V, P);
    ++g_launch_count;
}

// ---------------------------------------------------------------------------------------------
// starting point shifts      (sypha_solver_init.cpp:617-637)
// x, s hold x~, s~ ; sc->min_x / min_s hold their minima
// ---------------------------------------------------------------------------------------------
__global__ void k_start_shift1(IpmVecs V)
{
    __shared__ double sh[32];
    Scalars *sc = V.sc;
    const double dx = fmax(-1.5 * ord_decode(sc->min_x), 0.0);
    const double ds = fmax(-1.5 * ord_decode(sc->min_s), 0.0);
    double a0 = 0.0, a1 = 0.0, a2 = 0.0;
    for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < V.n; j += gridDim.x * blockDim.x)
    {
        const double xj = V.x[j] + dx, sj = V.s[j] + ds;
        V.x[j] = xj;
        V.s[j] = sj;
        a0 += xj * sj;
        a1 += xj;
        a2 += sj;
    }
    a0 = block_sum(a0, sh);
    a1 = block_sum(a1, sh);
    a2 = block_sum(a2, sh);
    if (threadIdx.x == 0)
    {
        V.partial[blockIdx.x] = a0;
        V.partial[SB200_MAX_PARTIAL_BLOCKS + blockIdx.x] = a1;
        V.partial[2 * SB200_MAX_PARTIAL_BLOCKS + blockIdx.x] = a2;
    }
    if (last_block_arrives(&sc->ticket[3], gridDim.x))
    {
        const double t0 = reduce_partials(V.partial, gridDim.x, sh);
        const double t1 = reduce_partials(V.partial + SB200_MAX_PARTIAL_BLOCKS, gridDim.x, sh);
        const double t2 = reduce_partials(V.partial + 2 * SB200_MAX_PARTIAL_BLOCKS, gridDim.x, sh);
        if (threadIdx.x == 0)
        {
            sc->sum0 = t0; sc->sum1 = t1; sc->sum2 = t2;
            sc->min_x = sc->min_s = SB200_ORD_DBL_MAX;
        }
    }
}
__global__ void k_start_shift2(IpmVecs V)
{
    const Scalars *sc = V.sc;
    const double prod = 0.5 * sc->sum0;
    const double dx = prod / sc->sum2;     // prod / sumS
    const double ds = prod / sc->sum1;     // prod / sumX
    for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < V.n; j += gridDim.x * blockDim.x)
    {
        V.x[j] += dx;
        V.s[j] += ds;
    }
}
void launch_start_shift1(const IpmVecs &V, cudaStream_t st)
{
    k_start_shift1<<<grid_for(V.n, kBlock), kBlock, 0, st>>>(V);
    ++g_launch_count;
}
void launch_start_shift2(const IpmVecs &V, cudaStream_t st)
{
    k_start_shift2<<<grid_for(V.n, kBlock), kBlock, 0, st>>>(V);
    ++g_launch_count;
}

} // namespace sb200
