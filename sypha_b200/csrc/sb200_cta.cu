// sb200_cta.cu - the whole Mehrotra solve of ONE LP by ONE thread block: the throughput form of the hot path.
//
// The latency-oriented path (sb200_api.cu: ~12 kernels per iteration, a data-flow factorisation that spreads one
// matrix over the whole GPU) is the right shape for ONE LP.  For the many small LPs of branch-and-bound
// (/root/reference/src/sypha_solver_bnb_driver.cpp:789-859 solves them one at a time) it is not: a CUPTI timeline of
// 32 node LPs in flight (profiles/r2_e_bnb_timeline_32_slots.log) shows the factorisation kernels stretched from 240 us
// to 1.9 ms each and holding 76 % of all kernel time, and neither more slots nor fewer CTAs per factorisation moves the
// nodes/s (profiles/r2_f_bnb_slots_sweep.log): 12 launches per iteration per LP and CTAs that wait on each other are
// the cost.  Here an LP is one kernel launch and one CTA: starting point, every iteration (normal-matrix assembly,
// Cholesky, two solves, four sparse products, the fused vector steps, the termination test) with block barriers only -
// no inter-CTA synchronisation, no host round trip, no launch per phase.  148 such CTAs run side by side.
//
//   assembly   : the same gather over the compact symbolic structure as k_assemble_normal16 (same summation order)
//   Cholesky   : left-looking by 64-column tile columns, 128 x 64 accumulator per step on the FP64 tensor pipe
//                (mma.sync m8n8k4), operands of step s+1 arriving by cp.async while step s multiplies; the 64 x 64
//                diagonal tile is factored and inverted in shared memory, the rows below take X = C W' (DMMA again)
//   solves     : block forward / backward substitution with the tile inverses, the vector in shared memory
//   products   : warp per row (CSR), 8 lanes per column (CSC) with the epilogues of k_spmv_csc
//   reductions : block-wide, fixed order (deterministic)
// The arithmetic is the reference's algorithm (sypha_solver.cpp:375-797, sypha_solver_init.cpp:543-652) exactly as
// the multi-kernel path computes it; tests/test_gpu_cta.py holds the two paths to the same iteration counts and
// objectives.
#include "sb200_cta.cuh"
#include "sb200_dmma.cuh"

namespace sb200 {

namespace {

constexpr int NT = 512, NW = NT / 32;
constexpr int CH = 128;                                  // rows of an accumulator chunk (two tiles)
constexpr int STAGE_A = CH * KP * 8;                     // 36864
constexpr int STAGE_B = TB * KP * 8;                     // 18432
constexpr int STAGE = STAGE_A + STAGE_B;                 // one pipeline stage: [A | B]
constexpr int NSTAGE = 3;
constexpr int SP = TB + 1;                               // row stride of the diagonal-tile scratch
constexpr int OFF_S = 2 * STAGE;                         // the tile being factored: aliases stage 2 (idle by then)
constexpr int OFF_W = NSTAGE * STAGE;                    // inverse of the current diagonal tile: lives across a tile column
constexpr int OFF_T = OFF_W + TB * SP * 8;               // 16 x 48 scratch of the block inversion
constexpr int TP = 49;
constexpr int OFF_RED = OFF_T + 16 * TP * 8;             // reduction scratch
constexpr int SMEM_BYTES = OFF_RED + 64 * 8;             // 206,?00
constexpr int ASM_RING = 384;                            // chunks of a warp's ring in the streamed assembly (6 KB)
constexpr int D_STAGE_MAX = OFF_W / 8;                   // doubles of a staged vector (below the persistent regions)
static_assert(TB * SP * 8 <= STAGE_A, "the tile scratch must fit the A part of a stage");
static_assert(SMEM_BYTES <= 227 * 1024, "shared memory budget");
__device__ __forceinline__ unsigned char *stage_a(unsigned char *smem, int s) { return smem + s * STAGE; }
__device__ __forceinline__ unsigned char *stage_b(unsigned char *smem, int s) { return smem + s * STAGE + STAGE_A; }

__device__ __forceinline__ void cp_async16(void *dst_smem, const void *src)
{
    const unsigned s = (unsigned)__cvta_generic_to_shared(dst_smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// block-wide sum / min in a fixed order; the result is returned in EVERY thread
__device__ __forceinline__ double cta_sum(double v, double *red)
{
    v = warp_sum(v);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    double t = 0.0;
#pragma unroll
    for (int i = 0; i < NW; ++i) t += red[i];
    return t;
}
__device__ __forceinline__ double cta_min(double v, double *red)
{
    v = warp_min(v);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    double t = DBL_MAX;
#pragma unroll
    for (int i = 0; i < NW; ++i) t = fmin(t, red[i]);
    return t;
}
__device__ __forceinline__ double group8_sum(double v)
{
    v += __shfl_xor_sync(0xffffffffu, v, 1);
    v += __shfl_xor_sync(0xffffffffu, v, 2);
    v += __shfl_xor_sync(0xffffffffu, v, 4);
    return v;
}

// ---- M = A diag(d) A' -----------------------------------------------------------------------------------------------
__device__ __forceinline__ double gather8(uint4 v, const double *d)
{
    return ((d[v.x & 0xffffu] + d[v.x >> 16]) + (d[v.y & 0xffffu] + d[v.y >> 16])) +
           ((d[v.z & 0xffffu] + d[v.z >> 16]) + (d[v.w & 0xffffu] + d[v.w >> 16]));
}

__device__ void cta_assemble(const CtaLp &L, const double *d, unsigned char *smem)
{
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const double *dd = d;
    __syncthreads();
    if (L.nd <= D_STAGE_MAX)
    {   // d staged in shared memory: every gather of the pass below is a shared-memory read
        double *ds = reinterpret_cast<double *>(smem);
        for (int i = tid; i < L.nd; i += NT) ds[i] = d[i];
        dd = ds;
        __syncthreads();
    }
    double *M = L.M;
    const int ld = L.ld;
    const size_t ring_off = ((size_t)L.nd * 8 + 15) & ~(size_t)15;
    if (dd != d && ring_off + (size_t)NW * ASM_RING * 16 <= (size_t)OFF_RED)
    {   // STREAMED form.  The chunk lists of a row's entries are ONE contiguous run of 16-byte chunks: the warp copies it
        // through a ring in shared memory with cp.async, up to ASM_RING chunks ahead of the entries it is summing, so the
        // DRAM latency of the lists (148 blocks x 38 MB never sit in L2) is paid once per row instead of once per step.
        // Same accumulators per entry as the loop below ((a0 + a1) + a2 by chunk position), so the same sums.
        uint4 *ring = reinterpret_cast<uint4 *>(smem + ring_off) + (size_t)w * ASM_RING;
        const double *dsm = reinterpret_cast<const double *>(smem);      // = dd, in a form the compiler knows is shared memory
        for (int i = 1 + w; i < L.base_m; i += NW)
        {
            const long long p0 = (long long)i * (i + 1) / 2;
            const unsigned int cs = __ldg(L.chunk_ptr + p0), ce = __ldg(L.chunk_ptr + p0 + i);
            unsigned int prod = cs;                                       // first chunk not yet requested
            unsigned int a_l = ce, e_l = ce;
            if (lane < i)
            {
                a_l = __ldg(L.chunk_ptr + p0 + lane);
                e_l = __ldg(L.chunk_ptr + p0 + lane + 1);
            }
            for (int k0 = 0; k0 < i; k0 += 32)
            {
                unsigned int a_n = ce, e_n = ce;                          // the next step's pointers, a step early
                if (k0 + 32 + lane < i)
                {
                    a_n = __ldg(L.chunk_ptr + p0 + k0 + 32 + lane);
                    e_n = __ldg(L.chunk_ptr + p0 + k0 + 33 + lane);
                }
                const unsigned int a_t = __shfl_sync(0xffffffffu, a_l, 0);
                const unsigned int b_t = __shfl_sync(0xffffffffu, e_l, min(31, i - k0 - 1));
                __syncwarp();
                const bool ahead = b_t <= prod;                           // this step's chunks were requested earlier
                const unsigned int lim = min(ce, a_t + (unsigned int)ASM_RING);
                for (unsigned int c = prod + lane; c < lim; c += 32) cp_async16(ring + (c - cs) % ASM_RING, L.term8 + c);
                if (lim > prod) prod = lim;
                cp_async_commit();
                if (ahead) cp_async_wait<1>();
                else cp_async_wait<0>();
                __syncwarp();
                double a0 = 0.0, a1 = 0.0, a2 = 0.0;
                if (b_t <= prod)
                {
                    for (unsigned int c = a_l; c < e_l; c += 3)
                    {
                        const bool h1 = c + 1 < e_l, h2 = c + 2 < e_l;
                        const uint4 v0 = ring[(c - cs) % ASM_RING];
                        a0 += gather8(v0, dsm);
                        if (h1) a1 += gather8(ring[(c + 1 - cs) % ASM_RING], dsm);
                        if (h2) a2 += gather8(ring[(c + 2 - cs) % ASM_RING], dsm);
                    }
                }
                else
                {   // a step whose lists exceed the ring: straight from global memory
                    for (unsigned int c = a_l; c < e_l; c += 3)
                    {
                        const bool h1 = c + 1 < e_l, h2 = c + 2 < e_l;
                        const uint4 z = make_uint4(0u, 0u, 0u, 0u);
                        const uint4 v0 = __ldg(L.term8 + c), v1 = h1 ? __ldg(L.term8 + c + 1) : z, v2 = h2 ? __ldg(L.term8 + c + 2) : z;
                        a0 += gather8(v0, dsm);
                        if (h1) a1 += gather8(v1, dsm);
                        if (h2) a2 += gather8(v2, dsm);
                    }
                }
                if (k0 + lane < i) M[(size_t)i * ld + k0 + lane] = (a0 + a1) + a2;
                a_l = a_n;
                e_l = e_n;
            }
            cp_async_wait<0>();
            __syncwarp();
        }
    }
    else
    // off-diagonal entries (i, k), k < i: a warp per row, lanes along the row (the entry list of a row is contiguous),
    // two entries per lane in flight (16 warps have little else to hide the chunk loads behind).  The accumulator a
    // chunk goes to depends on its position in the entry's list only: the same sums as k_assemble_normal16.
    for (int i = 1 + w; i < L.base_m; i += NW)
    {
        const long long p0 = (long long)i * (i + 1) / 2;
        for (int k = lane; k < i; k += 64)
        {
            const bool two = k + 32 < i;
            unsigned int ca = __ldg(L.chunk_ptr + p0 + k);
            const unsigned int ea = __ldg(L.chunk_ptr + p0 + k + 1);
            unsigned int cb = 0, eb = 0;
            if (two)
            {
                cb = __ldg(L.chunk_ptr + p0 + k + 32);
                eb = __ldg(L.chunk_ptr + p0 + k + 33);
            }
            double a0 = 0.0, a1 = 0.0, a2 = 0.0, b0 = 0.0, b1 = 0.0, b2 = 0.0;
            while (ca < ea || cb < eb)
            {
                const uint4 z = make_uint4(0u, 0u, 0u, 0u);
                const bool ha0 = ca < ea, ha1 = ca + 1 < ea, ha2 = ca + 2 < ea;
                const bool hb0 = cb < eb, hb1 = cb + 1 < eb, hb2 = cb + 2 < eb;
                const uint4 va0 = ha0 ? __ldg(L.term8 + ca) : z, va1 = ha1 ? __ldg(L.term8 + ca + 1) : z,
                            va2 = ha2 ? __ldg(L.term8 + ca + 2) : z;
                const uint4 vb0 = hb0 ? __ldg(L.term8 + cb) : z, vb1 = hb1 ? __ldg(L.term8 + cb + 1) : z,
                            vb2 = hb2 ? __ldg(L.term8 + cb + 2) : z;
                if (ha0) a0 += gather8(va0, dd);
                if (ha1) a1 += gather8(va1, dd);
                if (ha2) a2 += gather8(va2, dd);
                if (hb0) b0 += gather8(vb0, dd);
                if (hb1) b1 += gather8(vb1, dd);
                if (hb2) b2 += gather8(vb2, dd);
                ca += 3;
                cb += 3;
            }
            M[(size_t)i * ld + k] = (a0 + a1) + a2;
            if (two) M[(size_t)i * ld + k + 32] = (b0 + b1) + b2;
        }
    }
    for (int r = w; r < L.base_m; r += NW)
    {
        const long long pd = (long long)r * (r + 1) / 2 + r;
        const unsigned int da = __ldg(L.chunk_ptr + pd), de = __ldg(L.chunk_ptr + pd + 1);
        double sum = 0.0;
        for (unsigned int c = da + lane; c < de; c += 32) sum += gather8(__ldg(L.term8 + c), dd);
        sum = warp_sum(sum);
        if (lane == 0) M[(size_t)r * ld + r] = sum;
    }
    // rows of the node's branch decisions (k_assemble_extra_rows): row m0 + r = coef_r * column var_r of the base model
    for (int r = w; r < L.node_k; r += NW)
    {
        const int row = L.base_m + r, j = L.d_var[r];
        const double cf = L.d_coef[r], dj = d[j];
        double *Mr = M + (size_t)row * ld;
        for (int c = lane; c < row; c += 32) Mr[c] = 0.0;
        __syncwarp();
        for (int t = L.base_colptr[j] + lane; t < L.base_colptr[j + 1]; t += 32) Mr[L.base_rows[t]] = cf * L.base_cvals[t] * dj;
        for (int q = lane; q < r; q += 32)
            if (L.d_var[q] == j) Mr[L.base_m + q] = cf * L.d_coef[q] * dj;
        if (lane == 0) Mr[row] = cf * cf * dj + d[L.base_n + r];
    }
    __syncthreads();
}

// per-phase wall time of the last solve (ns, thread 0, %globaltimer), left in the LAST row of the trace buffer:
// [assembly, factorisation, solves, A v products, A' v products + epilogues, vector steps, starting point, whole LP];
// the row before it splits the factorisation: [accumulation steps, diagonal tiles, chunk epilogues]
__device__ __forceinline__ unsigned long long now_ns()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

// ---- Cholesky ---------------------------------------------------------------------------------------------------
// operands of accumulation step `step` of tile column j for the chunk of rows r0..: A = L[r0.., 64 kt + kc ..+32),
// B = L[64 j .., same columns)
__device__ __forceinline__ void potrf_issue(const CtaLp &L, unsigned char *smem, int stage, int step, int r0, int rows_valid, int c0)
{
    const int tid = threadIdx.x;
    const int kcol = (step >> 1) * TB + (step & 1) * KC;
    double(*As)[KP] = reinterpret_cast<double(*)[KP]>(stage_a(smem, stage));
    double(*Bs)[KP] = reinterpret_cast<double(*)[KP]>(stage_b(smem, stage));
    for (int idx = tid; idx < CH * (KC / 2); idx += NT)
    {
        const int r = idx >> 4, q = idx & 15;
        if (r < rows_valid) cp_async16(&As[r][2 * q], L.M + (size_t)(r0 + r) * L.ld + kcol + 2 * q);
    }
    for (int idx = tid; idx < TB * (KC / 2); idx += NT)
    {
        const int r = idx >> 4, q = idx & 15;
        cp_async16(&Bs[r][2 * q], L.M + (size_t)(c0 + r) * L.ld + kcol + 2 * q);
    }
    cp_async_commit();
}

// warp_mma over the first `kend` (multiple of 4) of the KC columns staged
template <int MI, int NJ>
__device__ __forceinline__ void warp_mma_k(const double (*As)[KP], const double (*Bs)[KP], int row0, int col0, int lane, int kend,
                                           double acc[MI][NJ][2])
{
    const int g = lane >> 2, tg = lane & 3;
    for (int kk = 0; kk < kend; kk += 4)
    {
        double a[MI], b[NJ];
#pragma unroll
        for (int i = 0; i < MI; ++i) a[i] = As[row0 + i * 8 + g][kk + tg];
#pragma unroll
        for (int j = 0; j < NJ; ++j) b[j] = Bs[col0 + j * 8 + g][kk + tg];
#pragma unroll
        for (int i = 0; i < MI; ++i)
#pragma unroll
            for (int j = 0; j < NJ; ++j) dmma_8x8x4(acc[i][j][0], acc[i][j][1], a[i], b[j]);
    }
}

// In-place Cholesky of the 64 x 64 tile in S (lower triangle), then Wm = S^-1 (lower; upper part zero).  Blocked by 16
// columns: the 16 x 16 diagonal block by one warp with rows in registers (shuffles carry the pivot row), the rows
// below by forward substitution (a thread per row), the trailing update by all threads - three barriers per panel
// instead of one per column (the first version, a rank-1 update per column, cost 58 us per tile: 27 % of the whole
// factorisation at m = 1000).  The inverse: the four diagonal blocks by a warp each, then block row by block row,
// W_ij = -W_ii sum_k L_ik W_kj.  *s_fail: 1-based index (within the tile) of the first non-positive pivot, 0 = none.
__device__ void tile_factor_invert(double (*S)[SP], double (*Wm)[SP], double (*Tm)[TP], int *s_fail)
{
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const unsigned full = 0xffffffffu;
    for (int b = 0; b < 4; ++b)
    {
        const int p0 = 16 * b;
        __syncthreads();
        if (w == 0)
        {
            const int rr = lane & 15, r = p0 + rr;
            double a[16];
#pragma unroll
            for (int c = 0; c < 16; ++c) a[c] = (c <= rr) ? S[r][p0 + c] : 0.0;
#pragma unroll
            for (int c = 0; c < 16; ++c)
            {
                const double piv = __shfl_sync(full, a[c], c);
                const bool bad = !(piv > 0.0);
                if (bad && lane == 0 && *s_fail == 0) *s_fail = p0 + c + 1;
                const double rs = bad ? 1.0 : rsqrt(piv);
                const double l = a[c] * rs;
                a[c] = l;
#pragma unroll
                for (int cc = c + 1; cc < 16; ++cc)
                {
                    const double lcc = __shfl_sync(full, l, cc);
                    a[cc] -= l * lcc;
                }
            }
            if (lane < 16)
            {
#pragma unroll
                for (int c = 0; c < 16; ++c)
                    if (c <= rr) S[r][p0 + c] = a[c];
            }
        }
        __syncthreads();
        const int rem = 48 - p0;                    // rows below the block
        if (tid < rem)
        {   // L[r][panel] = A[r][panel] D^-T: forward substitution along the row
            const int r = p0 + 16 + tid;
            double a[16];
#pragma unroll
            for (int c = 0; c < 16; ++c) a[c] = S[r][p0 + c];
#pragma unroll
            for (int c = 0; c < 16; ++c)
            {
                double v = a[c];
#pragma unroll
                for (int k = 0; k < c; ++k) v -= a[k] * S[p0 + c][p0 + k];
                a[c] = v / S[p0 + c][p0 + c];
            }
#pragma unroll
            for (int c = 0; c < 16; ++c) S[r][p0 + c] = a[c];
        }
        __syncthreads();
        for (int idx = tid; idx < rem * rem; idx += NT)
        {
            const int ri = idx / rem, ci = idx - ri * rem;
            if (ci <= ri)
            {
                const int r = p0 + 16 + ri, c = p0 + 16 + ci;
                double v = S[r][c];
#pragma unroll
                for (int k = 0; k < 16; ++k) v -= S[r][p0 + k] * S[c][p0 + k];
                S[r][c] = v;
            }
        }
    }
    __syncthreads();
    // ---- inverse ----------------------------------------------------------------------------------------------
    for (int idx = tid; idx < TB * TB; idx += NT)
    {
        const int r = idx >> 6, c = idx & 63;
        if ((c >> 4) > (r >> 4)) Wm[r][c] = 0.0;            // blocks above the block diagonal
    }
    if (w < 4 && lane < 16)
    {   // column `lane` of the inverse of diagonal block w
        const int p0 = 16 * w, c = lane;
        double x[16];
#pragma unroll
        for (int r = 0; r < 16; ++r)
        {
            double v = 0.0;
#pragma unroll
            for (int k = 0; k < r; ++k) v += S[p0 + r][p0 + k] * x[k];
            const double dinv = 1.0 / S[p0 + r][p0 + r];
            x[r] = (r < c) ? 0.0 : (r == c ? dinv : -v * dinv);
        }
#pragma unroll
        for (int r = 0; r < 16; ++r) Wm[p0 + r][p0 + c] = x[r];
    }
    __syncthreads();
    for (int i = 1; i < 4; ++i)
    {
        const int q0 = 16 * i;                      // block row i: columns 0 .. q0
        for (int idx = tid; idx < 16 * q0; idx += NT)
        {   // T = L[i, 0..i) W[0..i, 0..i)
            const int r = idx / q0, c = idx - r * q0;
            double v = 0.0;
            for (int k = c; k < q0; ++k) v += S[q0 + r][k] * Wm[k][c];
            Tm[r][c] = v;
        }
        __syncthreads();
        for (int idx = tid; idx < 16 * q0; idx += NT)
        {   // W[i, 0..i) = -W_ii T
            const int r = idx / q0, c = idx - r * q0;
            double v = 0.0;
            for (int q = 0; q <= r; ++q) v += Wm[q0 + r][q0 + q] * Tm[q][c];
            Wm[q0 + r][c] = -v;
        }
        __syncthreads();
    }
}

__device__ void cta_potrf(const CtaLp &L, unsigned char *smem, int *s_fail, int *info_out, double *sub)
{
    unsigned long long tq = now_ns();
#define SUBT(slot) do { if (threadIdx.x == 0) { const unsigned long long t__ = now_ns(); sub[slot] += (double)(t__ - tq); tq = t__; } } while (0)
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const int g = lane >> 2, tg = lane & 3;
    const int row0 = (w >> 2) * 32, col0 = (w & 3) * 16;         // 16 warps: 4 x 4, warp tile 32 x 16
    const int mpad = L.V.mpad, T = mpad / TB, ld = L.ld;
    double *M = L.M;
    double(*S)[SP] = reinterpret_cast<double(*)[SP]>(smem + OFF_S);
    double(*Wm)[SP] = reinterpret_cast<double(*)[SP]>(smem + OFF_W);
    double(*Tm)[TP] = reinterpret_cast<double(*)[TP]>(smem + OFF_T);
    if (tid == 0) *s_fail = 0;
    __syncthreads();
    for (int j = 0; j < T; ++j)
    {
        const int c0 = j * TB;
        for (int r0 = c0; r0 < mpad; r0 += CH)
        {
            const bool first = (r0 == c0);
            const int rows_valid = min(CH, mpad - r0);
            const bool wactive = row0 < rows_valid;
            double acc[4][2][2];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int jj = 0; jj < 2; ++jj)
                {
                    double2 v = make_double2(0.0, 0.0);
                    if (wactive)
                        v = *reinterpret_cast<const double2 *>(M + (size_t)(r0 + row0 + i * 8 + g) * ld + c0 + col0 + jj * 8 + tg * 2);
                    acc[i][jj][0] = v.x;
                    acc[i][jj][1] = v.y;
                }
            // three-stage pipeline, one barrier per step: the barrier of step s also says that every warp is done with
            // step s - 1, whose stage the loads of step s + 2 may then overwrite
            const int nsteps = 2 * j;
            if (nsteps > 0) potrf_issue(L, smem, 0, 0, r0, rows_valid, c0);
            if (nsteps > 1) potrf_issue(L, smem, 1, 1, r0, rows_valid, c0);
            for (int s = 0; s < nsteps; ++s)
            {
                if (s + 1 < nsteps)
                    cp_async_wait<1>();
                else
                    cp_async_wait<0>();
                __syncthreads();
                if (s + 2 < nsteps) potrf_issue(L, smem, (s + 2) % NSTAGE, s + 2, r0, rows_valid, c0);
                if (wactive)
                    warp_mma<4, 2>(reinterpret_cast<const double(*)[KP]>(stage_a(smem, s % NSTAGE)),
                                   reinterpret_cast<const double(*)[KP]>(stage_b(smem, s % NSTAGE)), row0, col0, lane, -1.0, acc);
            }
            __syncthreads();
            SUBT(0);
            // C = acc into the A parts of stages 0 and 1 (columns 0..31 | 32..63) for the product with W'
            if (wactive)
            {
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int jj = 0; jj < 2; ++jj)
                    {
                        const int r = row0 + i * 8 + g, c = col0 + jj * 8 + tg * 2;
                        double(*Cs)[KP] = reinterpret_cast<double(*)[KP]>(stage_a(smem, c >> 5));
                        Cs[r][c & 31] = acc[i][jj][0];
                        Cs[r][(c & 31) + 1] = acc[i][jj][1];
                    }
            }
            __syncthreads();
            SUBT(2);
            if (first)
            {
                for (int idx = tid; idx < TB * TB; idx += NT)
                {
                    const int r = idx >> 6, c = idx & 63;
                    if (c <= r) S[r][c] = reinterpret_cast<const double(*)[KP]>(stage_a(smem, c >> 5))[r][c & 31];
                }
                tile_factor_invert(S, Wm, Tm, s_fail);
                if (*s_fail && tid == 0 && *info_out == 0) *info_out = c0 + *s_fail;
                for (int idx = tid; idx < TB * TB; idx += NT)
                {
                    const int r = idx >> 6, c = idx & 63;
                    if (c <= r) M[(size_t)(c0 + r) * ld + c0 + c] = S[r][c];
                    L.linv[(size_t)j * TB * TB + idx] = Wm[r][c];
                }
                SUBT(1);
            }
            // W into the B parts of stages 0 and 1: Bs[h][c][k] = W[c][32 h + k]
            for (int idx = tid; idx < TB * TB; idx += NT)
            {
                const int c = idx >> 6, k = idx & 63;
                reinterpret_cast<double(*)[KP]>(stage_b(smem, k >> 5))[c][k & 31] = Wm[c][k];
            }
            __syncthreads();
            const bool xactive = wactive && !(first && row0 < TB);      // the diagonal tile itself is done
            if (xactive)
            {   // X = C W': W is lower triangular, so output columns col0 .. col0 + 15 only need k <= col0 + 15
                double x[4][2][2];
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int jj = 0; jj < 2; ++jj) x[i][jj][0] = x[i][jj][1] = 0.0;
                const int kmax = col0 + 16;
                warp_mma_k<4, 2>(reinterpret_cast<const double(*)[KP]>(stage_a(smem, 0)),
                                 reinterpret_cast<const double(*)[KP]>(stage_b(smem, 0)), row0, col0, lane, min(kmax, KC), x);
                if (kmax > KC)
                    warp_mma_k<4, 2>(reinterpret_cast<const double(*)[KP]>(stage_a(smem, 1)),
                                     reinterpret_cast<const double(*)[KP]>(stage_b(smem, 1)), row0, col0, lane, kmax - KC, x);
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int jj = 0; jj < 2; ++jj)
                        *reinterpret_cast<double2 *>(M + (size_t)(r0 + row0 + i * 8 + g) * ld + c0 + col0 + jj * 8 + tg * 2) =
                            make_double2(x[i][jj][0], x[i][jj][1]);
            }
            __syncthreads();
            SUBT(2);
        }
    }
#undef SUBT
}

// ---- L L' x = rhs, in place (rhs has mpad entries, zero beyond m) ---------------------------------------------------------
__device__ void cta_solve(const CtaLp &L, double *rhs, unsigned char *smem)
{
    const int tid = threadIdx.x;
    const int mpad = L.V.mpad, T = mpad / TB, ld = L.ld;
    const double *M = L.M;
    double *yv = reinterpret_cast<double *>(smem);                      // [mpad]
    double *tv = yv + CTA_MAX_MPAD;                                     // [64]
    double *part = tv + TB;                                             // [8][64]
    __syncthreads();
    for (int i = tid; i < mpad; i += NT) yv[i] = rhs[i];
    __syncthreads();
    {   // forward: y_j = W_j (b_j - sum_{k<j} L_jk y_k)
        const int r = tid >> 3, l = tid & 7;
        for (int j = 0; j < T; ++j)
        {
            const int row = j * TB + r;
            const double *Lr = M + (size_t)row * ld;
            double sum = 0.0;
            for (int k = l; k < j * TB; k += 8) sum += Lr[k] * yv[k];
            sum = group8_sum(sum);
            if (l == 0) tv[r] = yv[row] - sum;
            __syncthreads();
            const double *Wr = L.linv + (size_t)j * TB * TB + (size_t)r * TB;
            double s2 = 0.0;
            for (int k = l; k <= r; k += 8) s2 += Wr[k] * tv[k];
            s2 = group8_sum(s2);
            if (l == 0) yv[row] = s2;
            __syncthreads();
        }
    }
    {   // backward: x_j = W_j' (y_j - sum_{k>j} L_kj' x_k)
        const int c = tid & 63, l = tid >> 6;
        for (int j = T - 1; j >= 0; --j)
        {
            double sum = 0.0;
            for (int r = (j + 1) * TB + l; r < mpad; r += 8) sum += M[(size_t)r * ld + j * TB + c] * yv[r];
            part[l * TB + c] = sum;
            __syncthreads();
            if (tid < TB)
            {
                double s = 0.0;
#pragma unroll
                for (int q = 0; q < 8; ++q) s += part[q * TB + tid];
                tv[tid] = yv[j * TB + tid] - s;
            }
            __syncthreads();
            const double *Wj = L.linv + (size_t)j * TB * TB;
            double s2 = 0.0;
            for (int r = c + l; r < TB; r += 8) s2 += Wj[(size_t)r * TB + c] * tv[r];
            part[l * TB + c] = s2;
            __syncthreads();
            if (tid < TB)
            {
                double s = 0.0;
#pragma unroll
                for (int q = 0; q < 8; ++q) s += part[q * TB + tid];
                yv[j * TB + tid] = s;
            }
            __syncthreads();
        }
    }
    for (int i = tid; i < mpad; i += NT) rhs[i] = yv[i];
    __syncthreads();
}

// out[row] = alpha (A x)_row + beta z[row], rows < m: warp per row, four index/value pairs per lane in flight, x staged
// in shared memory when it fits (every gather is then a shared-memory read)
__device__ void cta_spmv_csr12(const CtaLp &L, const double *x, const double *z, double *out, double alpha, double beta,
                               unsigned char *smem)
{
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const double *xs = x;
    __syncthreads();
    if (L.V.n <= D_STAGE_MAX)
    {
        double *buf = reinterpret_cast<double *>(smem);
        for (int i = threadIdx.x; i < L.V.n; i += NT) buf[i] = x[i];
        xs = buf;
        __syncthreads();
    }
    for (int row = w; row < L.V.m; row += NW)
    {
        const int a = L.csr_offs[row], e = L.csr_offs[row + 1];
        double acc0 = 0.0, acc1 = 0.0, acc2 = 0.0, acc3 = 0.0;
        for (int k = a + lane; k < e; k += 128)
        {
            const int k1 = k + 32, k2 = k + 64, k3 = k + 96;
            const bool h1 = k1 < e, h2 = k2 < e, h3 = k3 < e;
            const int i0 = __ldg(L.csr_inds + k), i1 = h1 ? __ldg(L.csr_inds + k1) : 0, i2 = h2 ? __ldg(L.csr_inds + k2) : 0,
                      i3 = h3 ? __ldg(L.csr_inds + k3) : 0;
            const double v0 = L.csr_vals[k], v1 = h1 ? L.csr_vals[k1] : 0.0, v2 = h2 ? L.csr_vals[k2] : 0.0,
                         v3 = h3 ? L.csr_vals[k3] : 0.0;
            acc0 += v0 * xs[i0];
            acc1 += v1 * xs[i1];
            acc2 += v2 * xs[i2];
            acc3 += v3 * xs[i3];
        }
        double acc = (acc0 + acc1) + (acc2 + acc3);
        acc = warp_sum(acc);
        if (lane == 0) out[row] = (beta == 0.0) ? alpha * acc : alpha * acc + beta * z[row];
    }
    __syncthreads();
}

// w = A' v by 8 lanes per column (v staged in shared memory), with the epilogues of k_spmv_csc; returns the two minima
// in every thread
template <int MODE>
__device__ void cta_spmv_csc12(const CtaLp &L, const double *v, double *red, double *min0, double *min1, unsigned char *smem)
{
    const IpmVecs &V = L.V;
    const int tid = threadIdx.x, gl = tid & 7;
    const int n = V.n, nround = (n + 63) / 64 * 64;
    double *vs = reinterpret_cast<double *>(smem);                    // m <= mpad <= CTA_MAX_MPAD doubles
    __syncthreads();
    for (int i = tid; i < V.m; i += NT) vs[i] = v[i];
    __syncthreads();
    double m0 = DBL_MAX, m1 = DBL_MAX;
    for (int col = tid >> 3; col < nround; col += NT / 8)
    {
        double acc0 = 0.0, acc1 = 0.0;
        if (col < n)
        {
            const int a = L.csc_colptr[col], e = L.csc_colptr[col + 1];
            for (int k = a + gl; k < e; k += 16)
            {
                const int k1 = k + 8;
                const bool h1 = k1 < e;
                const int r0 = __ldg(L.csc_rows + k), r1 = h1 ? __ldg(L.csc_rows + k1) : 0;
                const double v0 = L.csc_vals[k], v1 = h1 ? L.csc_vals[k1] : 0.0;
                acc0 += v0 * vs[r0];
                acc1 += v1 * vs[r1];
            }
        }
        const double acc = group8_sum(acc0 + acc1);
        if (gl == 0 && col < n)
        {
            if (MODE == CSC_RECOVER)
            {
                const double ds = V.resC[col] - acc;
                const double xj = V.x[col], sj = V.s[col];
                const double dx = (V.resXS[col] - xj * ds) / sj;
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
    if (MODE != CSC_RESC)
    {
        *min0 = cta_min(m0, red);
        *min1 = cta_min(m1, red);
    }
    __syncthreads();
}

// ---- the same two products over the pattern-only lists of the base model (CompactLists): 2 bytes per entry instead of
//      12, whole 16-byte chunks per load, chunk pointers and vectors staged in shared memory so that every global load of
//      a row / column is independent of the others (the 12-byte forms above are chains of dependent loads per column).
//      The node's branch rows (coef x_var - slack) and their slack columns are added from d_var / d_coef. ----------------
__device__ void cta_av16(const CtaLp &L, const double *x, const double *z, double *out, double alpha, double beta,
                         unsigned char *smem)
{
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const int bm = L.base_m, bn = L.base_n;
    double *xs = reinterpret_cast<double *>(smem);                         // [bn + 1]: sign_j x_j, then the pad slot
    unsigned int *rp = reinterpret_cast<unsigned int *>(xs + bn + 1);      // [bm + 1]
    __syncthreads();
    for (int j = tid; j < bn; j += NT) xs[j] = L.col_sign[j] * x[j];
    if (tid == 0) xs[bn] = 0.0;
    for (int i = tid; i <= bm; i += NT) rp[i] = L.row_ptr[i];
    __syncthreads();
    const uint4 zero4 = make_uint4(0u, 0u, 0u, 0u);
    for (int r0 = w; r0 < bm; r0 += 2 * NW)
    {   // two rows per warp at a time, two chunks per lane and row in flight
        const int r1 = r0 + NW;
        const bool two = r1 < bm;
        unsigned int ca = rp[r0] + lane, cb = two ? rp[r1] + lane : 0u;
        const unsigned int ea = rp[r0 + 1], eb = two ? rp[r1 + 1] : 0u;
        double a0 = 0.0, a1 = 0.0, b0 = 0.0, b1 = 0.0;
        while (ca < ea || cb < eb)
        {
            const bool ha0 = ca < ea, ha1 = ca + 32 < ea, hb0 = cb < eb, hb1 = cb + 32 < eb;
            const uint4 ua0 = ha0 ? __ldg(L.row16 + ca) : zero4, ua1 = ha1 ? __ldg(L.row16 + ca + 32) : zero4;
            const uint4 ub0 = hb0 ? __ldg(L.row16 + cb) : zero4, ub1 = hb1 ? __ldg(L.row16 + cb + 32) : zero4;
            if (ha0) a0 += gather8(ua0, xs);
            if (ha1) a1 += gather8(ua1, xs);
            if (hb0) b0 += gather8(ub0, xs);
            if (hb1) b1 += gather8(ub1, xs);
            ca += 64;
            cb += 64;
        }
        const double sa = warp_sum(a0 + a1), sb = warp_sum(b0 + b1);
        if (lane == 0)
        {
            out[r0] = (beta == 0.0) ? alpha * sa : alpha * sa + beta * z[r0];
            if (two) out[r1] = (beta == 0.0) ? alpha * sb : alpha * sb + beta * z[r1];
        }
    }
    for (int r = tid; r < L.node_k; r += NT)
    {
        const double acc = L.d_coef[r] * x[L.d_var[r]] - x[bn + r];
        out[bm + r] = (beta == 0.0) ? alpha * acc : alpha * acc + beta * z[bm + r];
    }
    __syncthreads();
}

template <int MODE>
__device__ void cta_atv16(const CtaLp &L, const double *v, double *red, double *min0, double *min1, unsigned char *smem)
{
    const IpmVecs &V = L.V;
    const int tid = threadIdx.x, gl = tid & 7, g = tid >> 3;
    const int bm = L.base_m, bn = L.base_n, k = L.node_k, n = V.n;
    double *vs = reinterpret_cast<double *>(smem);                         // [bm + 1 + k]: v of the base rows, pad, node rows
    double *accs = vs + bm + 1 + k;                                        // [bn + k]: (A' v)_j
    unsigned int *cp = reinterpret_cast<unsigned int *>(accs + bn + k);    // [bn + 1]
    __syncthreads();
    for (int i = tid; i < bm; i += NT) vs[i] = v[i];
    if (tid == 0) vs[bm] = 0.0;
    for (int r = tid; r < k; r += NT) vs[bm + 1 + r] = v[bm + r];
    for (int j = tid; j <= bn; j += NT) cp[j] = L.col_ptr[j];
    __syncthreads();
    constexpr int G = NT / 8;
    const uint4 zero4 = make_uint4(0u, 0u, 0u, 0u);
    for (int cw = g & ~3; cw < bn; cw += 4 * G)
    {   // 8 lanes per column, four columns per group in flight; the trip count is the warp's (the shuffles below need
        // all 32 lanes), columns beyond the last are empty lists
        const int c0 = cw + (g & 3);
        unsigned int a[4], e[4];
        double acc[4];
#pragma unroll
        for (int q = 0; q < 4; ++q)
        {
            const int cq = c0 + q * G;
            const bool in = cq < bn;
            a[q] = in ? cp[cq] + gl : 0u;
            e[q] = in ? cp[cq + 1] : 0u;
            acc[q] = 0.0;
        }
        while (a[0] < e[0] || a[1] < e[1] || a[2] < e[2] || a[3] < e[3])
        {
            uint4 u[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) u[q] = a[q] < e[q] ? __ldg(L.col16 + a[q]) : zero4;
#pragma unroll
            for (int q = 0; q < 4; ++q)
            {
                if (a[q] < e[q]) acc[q] += gather8(u[q], vs);
                a[q] += 8;
            }
        }
#pragma unroll
        for (int q = 0; q < 4; ++q)
        {
            const double t = group8_sum(acc[q]);
            const int cq = c0 + q * G;
            if (gl == 0 && cq < bn) accs[cq] = t;
        }
    }
    __syncthreads();
    for (int j = tid; j < bn; j += NT) accs[j] *= L.col_sign[j];
    for (int r = tid; r < k; r += NT) accs[bn + r] = -vs[bm + 1 + r];
    if (k)
    {
        __syncthreads();
        if (tid == 0)
            for (int r = 0; r < k; ++r) accs[L.d_var[r]] += L.d_coef[r] * vs[bm + 1 + r];
    }
    __syncthreads();
    double m0 = DBL_MAX, m1 = DBL_MAX;
#pragma unroll 2
    for (int col = tid; col < n; col += NT)
    {
        const double acc = accs[col];
        if (MODE == CSC_RECOVER)
        {
            const double ds = V.resC[col] - acc;
            const double xj = V.x[col], sj = V.s[col];
            const double dx = (V.resXS[col] - xj * ds) / sj;
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
    if (MODE != CSC_RESC)
    {
        *min0 = cta_min(m0, red);
        *min1 = cta_min(m1, red);
    }
    __syncthreads();
}

__device__ __forceinline__ void cta_spmv_csr(const CtaLp &L, const double *x, const double *z, double *out, double alpha,
                                             double beta, unsigned char *smem)
{
    if (L.row16) cta_av16(L, x, z, out, alpha, beta, smem);
    else cta_spmv_csr12(L, x, z, out, alpha, beta, smem);
}
template <int MODE>
__device__ __forceinline__ void cta_spmv_csc(const CtaLp &L, const double *v, double *red, double *min0, double *min1,
                                             unsigned char *smem)
{
    if (L.row16) cta_atv16<MODE>(L, v, red, min0, min1, smem);
    else cta_spmv_csc12<MODE>(L, v, red, min0, min1, smem);
}

__device__ __forceinline__ void prologue_elem(const IpmVecs &V, int j, double xj, double sj, double rc)
{
    const double rxs = -xj * sj;
    V.resXS[j] = rxs;
    V.d[j] = xj / sj;
    V.t[j] = (xj * rc - rxs) / sj;
}

#define PH(slot) do { if (tid == 0) { const unsigned long long t__ = now_ns(); ph[slot] += (double)(t__ - t_last); t_last = t__; } } while (0)

__global__ void __launch_bounds__(NT, 1) k_ipm_cta(const CtaLp *lps)
{
    extern __shared__ __align__(16) unsigned char smem[];
    __shared__ CtaLp L;
    __shared__ int s_fail, s_info, s_done;
    const int tid = threadIdx.x;
    if (tid == 0) L = lps[blockIdx.x];
    __syncthreads();
    const IpmVecs &V = L.V;
    const DevParams P = *L.P;
    Scalars *sc = V.sc;
    double *red = reinterpret_cast<double *>(smem + OFF_RED);
    const int n = V.n, m = V.m, mpad = V.mpad;
    double mn0, mn1;
    double ph[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    double sub[3] = {0, 0, 0};
    unsigned long long t_last = now_ns();
    const unsigned long long t_begin = t_last;

    if (tid == 0) s_info = 0;
    __syncthreads();
    if (L.warm)
    {   // ---- warm start from the parent node's iterate: its optimal face pulled back into the interior ------------------
        const double *wx = L.warm, *wy = L.warm + L.warm_n, *ws_ = L.warm + L.warm_n + L.warm_m;
        const double fl = L.warm_floor;
        for (int j = tid; j < n; j += NT)
        {
            V.x[j] = j < L.warm_n ? fmax(wx[j], fl) : fl;
            V.s[j] = j < L.warm_n ? fmax(ws_[j], fl) : fl;
        }
        for (int i = tid; i < mpad; i += NT)
        {
            if (i < m) V.y[i] = i < L.warm_m ? wy[i] : 0.0;
            V.rhs[i] = 0.0;
        }
        __syncthreads();
    }
    else
    {
    // ---- starting point (sypha_solver_init.cpp:543-652): D = I ------------------------------------------------------
        cta_assemble(L, L.ones, smem);
        cta_potrf(L, smem, &s_fail, &s_info, sub);
        for (int i = tid; i < mpad; i += NT) V.rhs[i] = i < m ? V.b[i] : 0.0;
        cta_solve(L, V.rhs, smem);                                             // (A A')^-1 b
        cta_spmv_csc<CSC_START_X>(L, V.rhs, red, &mn0, &mn1, smem);                  // x~ = A' (.)
        const double min_x = mn0;
        cta_spmv_csr(L, V.c, nullptr, V.rhs, 1.0, 0.0, smem);                        // A c
        cta_solve(L, V.rhs, smem);                                             // y~
        for (int i = tid; i < m; i += NT) V.y[i] = V.rhs[i];
        __syncthreads();
        cta_spmv_csc<CSC_START_S>(L, V.y, red, &mn0, &mn1, smem);                    // s~ = c - A' y~
        const double min_s = mn1;
        {
            const double dx = fmax(-1.5 * min_x, 0.0), ds = fmax(-1.5 * min_s, 0.0);
            double a0 = 0.0, a1 = 0.0, a2 = 0.0;
            for (int j = tid; j < n; j += NT)
            {
                const double xj = V.x[j] + dx, sj = V.s[j] + ds;
                V.x[j] = xj;
                V.s[j] = sj;
                a0 += xj * sj;
                a1 += xj;
                a2 += sj;
            }
            a0 = cta_sum(a0, red);
            a1 = cta_sum(a1, red);
            a2 = cta_sum(a2, red);
            const double prod = 0.5 * a0, dx2 = prod / a2, ds2 = prod / a1;
            for (int j = tid; j < n; j += NT)
            {
                V.x[j] += dx2;
                V.s[j] += ds2;
            }
            __syncthreads();
        }
    }
    // ---- initial residuals and mu (sypha_solver.cpp:375-459) ----------------------------------------------------------
    cta_spmv_csc<CSC_RESC>(L, V.y, red, &mn0, &mn1, smem);                       // resC = c - s - A'y
    cta_spmv_csr(L, V.x, V.b, V.resB, -1.0, 1.0, smem);                          // resB = b - A x
    double mu, primal, dual;
    {
        double acc = 0.0, a_p = 0.0, a_d = 0.0;
        for (int j = tid; j < n; j += NT)
        {
            acc += V.x[j] * V.s[j];
            if (j < P.n_orig) a_p += V.x[j] * V.c[j];
        }
        for (int j = tid; j < m; j += NT) a_d += V.y[j] * V.b[j];
        mu = cta_sum(acc, red) / (double)n;
        primal = cta_sum(a_p, red);
        dual = cta_sum(a_d, red);
    }
    int iter = 0, stall = 0, done = 0, reason = SB200_TERM_MAX_ITER, numerical = 0;
    double best_gap = INFINITY, alpha_p = 0.0, alpha_d = 0.0, mu_aff = 0.0, sigma = 0.0;
    if (!(mu > P.mu_tol) || P.max_iter <= 0)
    {
        done = 1;
        reason = (mu <= P.mu_tol) ? SB200_TERM_CONVERGED : SB200_TERM_MAX_ITER;
    }
    if (s_info)
    {
        done = 1;
        numerical = 1;
        reason = SB200_TERM_INFEASIBLE_OR_NUMERICAL;
    }
    for (int j = tid; j < n; j += NT) prologue_elem(V, j, V.x[j], V.s[j], V.resC[j]);
    __syncthreads();
    PH(6);

    // ---- predictor-corrector loop (sypha_solver.cpp:496-772) ------------------------------------------------------------
    while (!done)
    {
        cta_assemble(L, V.d, smem);
        PH(0);
        cta_potrf(L, smem, &s_fail, &s_info, sub);
        PH(1);
        if (s_info)
        {   // non-positive pivot: the LP is reported infeasible-or-numerical (the reference's LU info != 0, :524-529)
            numerical = 1;
            reason = SB200_TERM_INFEASIBLE_OR_NUMERICAL;
            break;
        }
        cta_spmv_csr(L, V.t, V.resB, V.rhs, 1.0, 1.0, smem);                     // rhs = resB + A t
        PH(3);
        cta_solve(L, V.rhs, smem);                                         // dy (affine)
        PH(2);
        cta_spmv_csc<CSC_RECOVER>(L, V.rhs, red, &mn0, &mn1, smem);
        PH(4);
        {   // affine step lengths, mu_aff, sigma, corrector right-hand side (:596-629)
            const double ap = fmin(1.0, mn0), ad = fmin(1.0, mn1);
            double acc = 0.0;
            for (int j = tid; j < n; j += NT) acc += (V.x[j] + ap * V.dx[j]) * (V.s[j] + ad * V.ds[j]);
            mu_aff = cta_sum(acc, red) / (double)n;
            const double r = mu_aff / mu;
            sigma = r * r * r;
            const double sm = sigma * mu;
            for (int j = tid; j < n; j += NT)
            {
                const double corr = -V.dx[j] * V.ds[j] + sm;
                const double rxs = V.resXS[j] + corr;
                V.resXS[j] = rxs;
                V.t[j] = (V.x[j] * V.resC[j] - rxs) / V.s[j];
            }
            __syncthreads();
        }
        PH(5);
        cta_spmv_csr(L, V.t, V.resB, V.rhs, 1.0, 1.0, smem);
        PH(3);
        cta_solve(L, V.rhs, smem);                                         // dy (corrector)
        PH(2);
        cta_spmv_csc<CSC_RECOVER>(L, V.rhs, red, &mn0, &mn1, smem);
        PH(4);
        {   // step, residual scaling, mu / objectives / termination, next prologue (:693-769)
            const double ap = fmin(1.0, P.eta * mn0), ad = fmin(1.0, P.eta * mn1);
            const double fc = -(ad - 1.0), fb = -(ap - 1.0);
            double a_xs = 0.0, a_p = 0.0, a_d = 0.0;
            for (int j = tid; j < n; j += NT)
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
            for (int j = tid; j < m; j += NT)
            {
                const double yj = V.y[j] + ad * V.rhs[j];
                V.y[j] = yj;
                V.resB[j] *= fb;
                a_d += yj * V.b[j];
            }
            const double mu_in = mu;
            mu = cta_sum(a_xs, red) / (double)n;
            primal = cta_sum(a_p, red);
            dual = cta_sum(a_d, red);
            const double gap = fabs(primal - dual) / fmax(1.0, fabs(primal));
            alpha_p = ap;
            alpha_d = ad;
            if (tid == 0 && iter < SB200_TRACE_ROWS)
            {
                double *tr = V.trace + (size_t)iter * SB200_TRACE_COLS;
                tr[0] = mu_in; tr[1] = mu; tr[2] = mu_aff; tr[3] = sigma;
                tr[4] = ap; tr[5] = ad; tr[6] = primal; tr[7] = dual;
            }
            if (!isfinite(mu) || mu < 0.0 || !isfinite(primal) || !isfinite(dual) || !isfinite(gap))
            {   // :723-729, :748-753 (iteration not counted)
                numerical = 1;
                reason = SB200_TERM_INFEASIBLE_OR_NUMERICAL;
                done = 1;
            }
            else
            {
                if (gap < best_gap * (1.0 - P.min_improv_ratio))
                {
                    best_gap = gap;
                    stall = 0;
                }
                else if (P.gap_enabled)
                {
                    if (++stall >= P.gap_window)
                    {
                        reason = SB200_TERM_GAP_STALLED;
                        done = 1;
                    }
                }
                ++iter;
                if (!done && (iter >= P.max_iter || !(mu > P.mu_tol)))
                {
                    reason = (mu <= P.mu_tol) ? SB200_TERM_CONVERGED : SB200_TERM_MAX_ITER;
                    done = 1;
                }
            }
            __syncthreads();
        }
        PH(5);
    }
    // the node's final x | y | s for its children (sb200_node_delta.export_xys), written by the kernel itself
    if (L.export_xys)
    {
        double *ex = L.export_xys, *ey = ex + n, *es = ey + m;
        for (int j = tid; j < n; j += NT)
        {
            ex[j] = V.x[j];
            es[j] = V.s[j];
        }
        for (int i = tid; i < m; i += NT) ey[i] = V.y[i];
    }
    __threadfence();
    __syncthreads();
    if (tid == 0)
    {
        ph[7] = (double)(now_ns() - t_begin);
        double *tr = V.trace + (size_t)(SB200_TRACE_ROWS - 1) * SB200_TRACE_COLS;
#pragma unroll
        for (int q = 0; q < 8; ++q) tr[q] = ph[q];
        tr[-8] = sub[0];
        tr[-7] = sub[1];
        tr[-6] = sub[2];
        // the scalar block twice: in device memory (the other kernels of the library read it there) and in the pinned
        // host mirror, whose `done` flag - written last, behind a system-scope fence - is what the host polls: no event,
        // no copy, no stream synchronisation per node LP
        Scalars *dst[2] = {sc, L.sc_pinned};
        for (int q = 0; q < 2; ++q)
        {
            Scalars *o = dst[q];
            if (!o) continue;
            o->mu = mu;
            o->mu_aff = mu_aff;
            o->sigma = sigma;
            o->alpha_p = alpha_p;
            o->alpha_d = alpha_d;
            o->primal = primal;
            o->dual = dual;
            o->gap = fabs(primal - dual) / fmax(1.0, fabs(primal));
            o->best_gap = best_gap;
            o->iter = iter;
            o->stall = stall;
            o->reason = reason;
            o->numerical = numerical;
            o->chol_info = s_info;
            o->cg_total = 0;
            o->sum0 = ph[7];                     // the LP's own wall time in ns (%globaltimer)
        }
        sc->done = 1;
        if (L.sc_pinned)
        {
            __threadfence_system();
            *reinterpret_cast<volatile int *>(&L.sc_pinned->done) = 1;
        }
        (void)s_done;
    }
}

} // namespace

int cta_lp_smem_bytes() { return SMEM_BYTES; }

bool cta_lists_fit(int base_m, int base_n, int node_k)
{
    const size_t av = 8 * ((size_t)base_n + 1) + 4 * ((size_t)base_m + 1);
    const size_t atv = 8 * ((size_t)base_m + 1 + node_k) + 8 * ((size_t)base_n + node_k) + 4 * ((size_t)base_n + 1);
    return av <= (size_t)OFF_W && atv <= (size_t)OFF_W;
}

int launch_ipm_cta(const CtaLp *lps, int count, cudaStream_t st)
{
    static unsigned long long attr_seen = 0;
    if (first_use_on_device(attr_seen) &&
        cudaFuncSetAttribute(k_ipm_cta, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES) != cudaSuccess)
        return SB200_ERR_CUDA;
    k_ipm_cta<<<count, NT, SMEM_BYTES, st>>>(lps);
    ++g_launch_count;
    return SB200_OK;
}

} // namespace sb200
