// sb200_spmv.cu - CSR (A v) and CSC (A' v) sparse matrix-vector products with fused epilogues.
//
// Replaces the six cusparseSpMV call sites of the reference (SURVEY.md 2.2:
// /root/reference/src/sypha_solver.cpp:419,450; sypha_solver_krylov.cu:215,309,324,428) and the
// vector kernels around them (krylov.cu:26-43 jacobi_diag, :65-71 scale_by_D2, :74-82 recover_dx,
// sypha_solver_utils.cu:68-137 ratio test).  The reference multiplies by A' through
// CUSPARSE_OPERATION_TRANSPOSE on the CSR arrays; here a CSC copy (built once per model) makes the
// transposed product a gather with coalesced index/value streams and no atomics.
//
// Both kernels are HBM-bound: 12 B per stored entry (8 B value + 4 B index) + the dense vectors.
#include "sb200_kernels.cuh"
#include "sb200_pcg.cuh"

namespace sb200 {

// ---------------------------------------------------------------------------------------------
// CSR: one warp per row, lanes stride the row (coalesced inds/vals), shuffle reduction.
// ---------------------------------------------------------------------------------------------
template <int MODE>   // 0: out = alpha*Ax + beta*z ; 1: jacobi diag (x = d): sum a^2 d[col]
__global__ void __launch_bounds__(256)
k_spmv_csr(int m, const int *__restrict__ offs, const int *__restrict__ inds,
           const double *__restrict__ vals, const double *__restrict__ x,
           const double *__restrict__ z, double *__restrict__ out, double alpha, double beta)
{
    pdl_wait();
    const int lane = threadIdx.x & 31;
    const int wpb = blockDim.x >> 5;
    for (int row = blockIdx.x * wpb + (threadIdx.x >> 5); row < m; row += gridDim.x * wpb)
    {
        const int a = offs[row], e = offs[row + 1];
        // four index/value pairs per lane in flight: at OR-Library sizes (1000 rows of 500 entries) the
        // kernel is one warp per row on fewer warps than the GPU has schedulers, i.e. a chain of dependent
        // L2 latencies (index -> gather); unrolling the chain is what shortens it
        double acc0 = 0.0, acc1 = 0.0, acc2 = 0.0, acc3 = 0.0;
        for (int k = a + lane; k < e; k += 128)
        {
            const int k1 = k + 32, k2 = k + 64, k3 = k + 96;
            const bool h1 = k1 < e, h2 = k2 < e, h3 = k3 < e;
            const int i0 = inds[k], i1 = h1 ? inds[k1] : 0, i2 = h2 ? inds[k2] : 0, i3 = h3 ? inds[k3] : 0;
            const double v0 = vals[k], v1 = h1 ? vals[k1] : 0.0, v2 = h2 ? vals[k2] : 0.0, v3 = h3 ? vals[k3] : 0.0;
            const double x0 = __ldg(x + i0), x1 = __ldg(x + i1), x2 = __ldg(x + i2), x3 = __ldg(x + i3);
            acc0 += (MODE == 1) ? v0 * v0 * x0 : v0 * x0;
            acc1 += (MODE == 1) ? v1 * v1 * x1 : v1 * x1;
            acc2 += (MODE == 1) ? v2 * v2 * x2 : v2 * x2;
            acc3 += (MODE == 1) ? v3 * v3 * x3 : v3 * x3;
        }
        double acc = (acc0 + acc1) + (acc2 + acc3);
        acc = warp_sum(acc);
        if (lane == 0)
        {
            if (MODE == 1)
                out[row] = acc;
            else
                out[row] = (beta == 0.0) ? alpha * acc : alpha * acc + beta * z[row];
        }
    }
}

void launch_spmv_csr(const CsrView &A, const double *x, const double *z, double *out, double alpha,
                     double beta, cudaStream_t st)
{
    if (A.blk) return launch_blk_spmv_rows(*A.blk, x, z, out, alpha, beta, st);
    const int grid = grid_for((long long)A.m * 32, 256, 148 * 16);
    launch_pdl(k_spmv_csr<0>, grid, 256, 0, st, A.m, A.offs, A.inds, A.vals, x, z, out, alpha, beta);
    ++g_launch_count;
}
void launch_jacobi_diag(const CsrView &A, const double *d, double *diag, cudaStream_t st)
{
    if (A.blk) return launch_blk_jacobi_diag(*A.blk, d, diag, st);
    const int grid = grid_for((long long)A.m * 32, 256, 148 * 16);
    k_spmv_csr<1><<<grid, 256, 0, st>>>(A.m, A.offs, A.inds, A.vals, d, nullptr, diag, 1.0, 0.0);
    ++g_launch_count;
}

// ---------------------------------------------------------------------------------------------
// CSC: G lanes per column (G = 1..32 chosen from the mean column length), segmented shuffle
// reduction, epilogue by the group leader.
// ---------------------------------------------------------------------------------------------
// atomicMin with a relaxed pre-check: the value read can only be >= the current minimum, so skipping when
// key >= read never loses a smaller key; the result stays the exact minimum
__device__ __forceinline__ void ord_min(unsigned long long *addr, unsigned long long key)
{
    if (key < *reinterpret_cast<volatile unsigned long long *>(addr)) atomicMin(addr, key);
}

template <int G>
__device__ __forceinline__ double group_sum(double v)
{
#pragma unroll
    for (int o = G / 2; o > 0; o >>= 1)
        v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

template <int G, int MODE>
__global__ void __launch_bounds__(256)
k_spmv_csc(int n, const int *__restrict__ colptr, const int *__restrict__ rows,
           const double *__restrict__ vals, const double *__restrict__ v,
           const double *__restrict__ z, double *__restrict__ out, double alpha, double beta,
           IpmVecs V)
{
    __shared__ double sh[32];
    pdl_wait();
    if (MODE == CSC_RECOVER)
        if (V.sc->done) return;
    const int gl = threadIdx.x & (G - 1);
    const int groups_per_block = blockDim.x / G;
    const int ncols_round = ((n + groups_per_block - 1) / groups_per_block) * groups_per_block;
    double m0 = DBL_MAX, m1 = DBL_MAX;
    for (int col = blockIdx.x * groups_per_block + threadIdx.x / G; col < ncols_round;
         col += gridDim.x * groups_per_block)
    {
        double acc = 0.0;
        double e_rc = 0.0, e_x = 0.0, e_s = 1.0, e_rxs = 0.0;
        if (col < n)
        {
            const int a = colptr[col], e = colptr[col + 1];
            if (MODE == CSC_RECOVER && gl == 0)
            {   // epilogue operands requested together with the column, not after the reduction
                e_rc = V.resC[col];
                e_x = V.x[col];
                e_s = V.s[col];
                e_rxs = V.resXS[col];
            }
            for (int k = a + gl; k < e; k += G)
                acc += vals[k] * __ldg(v + rows[k]);
        }
        acc = group_sum<G>(acc);
        if (gl == 0 && col < n)
        {
            if (MODE == CSC_PLAIN)
                out[col] = (beta == 0.0) ? alpha * acc : alpha * acc + beta * z[col];
            else if (MODE == CSC_RECOVER)
            {   // krylov.cu:74-82 + utils.cu:68-79
                const double ds = e_rc - acc;
                const double xj = e_x, sj = e_s;
                const double dx = (e_rxs - xj * ds) / sj;
                V.ds[col] = ds;
                V.dx[col] = dx;
                if (dx < 0.0) m0 = fmin(m0, -xj / dx);
                if (ds < 0.0) m1 = fmin(m1, -sj / ds);
            }
            else if (MODE == CSC_START_X)
            {
                V.x[col] = acc;
                m0 = fmin(m0, acc);
            }
            else if (MODE == CSC_START_S)
            {
                const double sj = V.c[col] - acc;
                V.s[col] = sj;
                m1 = fmin(m1, sj);
            }
            else if (MODE == CSC_RESC)
                V.resC[col] = V.c[col] - V.s[col] - acc;
        }
    }
    if (MODE == CSC_RECOVER || MODE == CSC_START_X || MODE == CSC_START_S)
    {
        m0 = block_min(m0, sh);
        m1 = block_min(m1, sh);
        if (threadIdx.x == 0)
        {   // a CTA that cannot lower the minimum does not touch it (hundreds of CTAs, two addresses)
            if (MODE == CSC_RECOVER)
            {
                ord_min(&V.sc->amax_p, ord_encode(m0));
                ord_min(&V.sc->amax_d, ord_encode(m1));
            }
            else if (MODE == CSC_START_X)
                ord_min(&V.sc->min_x, ord_encode(m0));
            else
                ord_min(&V.sc->min_s, ord_encode(m1));
        }
    }
}

int pick_csc_lanes(long long nnz, int n)
{
    const double avg = n > 0 ? (double)nnz / n : 1.0;
    int g = 1;
    while (g < 32 && g * 2 <= avg) g <<= 1;    // largest power of two <= mean column length
    return g;
}

template <int G>
static void launch_csc_g(const CscView &A, int mode, const double *v, const double *z, double *out,
                         double alpha, double beta, const IpmVecs &V, cudaStream_t st)
{
    const int grid = grid_for((long long)A.n * G, 256, 148 * 16);
#define SB200_CSC_CASE(M)                                                                          \
    case M:                                                                                        \
        launch_pdl(k_spmv_csc<G, M>, grid, 256, 0, st, A.n, A.colptr, A.rows, A.vals, v, z, out,  \
                   alpha, beta, V);                                                               \
        break;
    switch (mode)
    {
        SB200_CSC_CASE(CSC_PLAIN)
        SB200_CSC_CASE(CSC_RECOVER)
        SB200_CSC_CASE(CSC_START_X)
        SB200_CSC_CASE(CSC_START_S)
        SB200_CSC_CASE(CSC_RESC)
    }
#undef SB200_CSC_CASE
    ++g_launch_count;
}

void launch_spmv_csc(const CscView &A, int mode, const double *v, const double *z, double *out,
                     double alpha, double beta, const IpmVecs *Vp, cudaStream_t st)
{
    if (A.blk) return launch_blk_spmv_cols(*A.blk, mode, v, z, out, alpha, beta, Vp, st);
    IpmVecs V{};
    if (Vp) V = *Vp;
    switch (A.lanes)
    {
    case 1: launch_csc_g<1>(A, mode, v, z, out, alpha, beta, V, st); break;
    case 2: launch_csc_g<2>(A, mode, v, z, out, alpha, beta, V, st); break;
    case 4: launch_csc_g<4>(A, mode, v, z, out, alpha, beta, V, st); break;
    case 8: launch_csc_g<8>(A, mode, v, z, out, alpha, beta, V, st); break;
    case 16: launch_csc_g<16>(A, mode, v, z, out, alpha, beta, V, st); break;
    default: launch_csc_g<32>(A, mode, v, z, out, alpha, beta, V, st); break;
    }
}

// PCG: q = D A'p (D = I when dscale == nullptr); skipped once the CG solve has finished
template <int G>
__global__ void __launch_bounds__(256)
k_spmv_csc_cg(int n, const int *__restrict__ colptr, const int *__restrict__ rows,
              const double *__restrict__ vals, const double *__restrict__ p, double *__restrict__ q,
              const double *__restrict__ dscale, const Scalars *sc)
{
    if (sc->cg_done) return;
    const int gl = threadIdx.x & (G - 1);
    const int groups_per_block = blockDim.x / G;
    const int ncols_round = ((n + groups_per_block - 1) / groups_per_block) * groups_per_block;
    for (int col = blockIdx.x * groups_per_block + threadIdx.x / G; col < ncols_round;
         col += gridDim.x * groups_per_block)
    {
        double acc = 0.0;
        if (col < n)
        {
            const int a = colptr[col], e = colptr[col + 1];
            for (int k = a + gl; k < e; k += G)
                acc += vals[k] * __ldg(p + rows[k]);
        }
        acc = group_sum<G>(acc);
        if (gl == 0 && col < n)
            q[col] = dscale ? dscale[col] * acc : acc;
    }
}

void launch_spmv_csc_cg(const CscView &A, const double *p, double *q, const double *dscale,
                        const Scalars *sc, cudaStream_t st)
{
    if (A.blk) return launch_blk_cg_cols(*A.blk, p, q, dscale, sc, st);
#define SB200_CG_CASE(G)                                                                           \
    case G:                                                                                        \
        k_spmv_csc_cg<G><<<grid_for((long long)A.n * G, 256, 148 * 16), 256, 0, st>>>(             \
            A.n, A.colptr, A.rows, A.vals, p, q, dscale, sc);                                      \
        break;
    switch (A.lanes)
    {
        SB200_CG_CASE(1)
        SB200_CG_CASE(2)
        SB200_CG_CASE(4)
        SB200_CG_CASE(8)
        SB200_CG_CASE(16)
    default:
        k_spmv_csc_cg<32><<<grid_for((long long)A.n * 32, 256, 148 * 16), 256, 0, st>>>(
            A.n, A.colptr, A.rows, A.vals, p, q, dscale, sc);
    }
#undef SB200_CG_CASE
    ++g_launch_count;
}

} // namespace sb200
