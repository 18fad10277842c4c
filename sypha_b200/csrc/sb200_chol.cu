// sb200_chol.cu - dense FP64 Cholesky of the normal matrix and the triangular solves.
//
// Replaces the reference's per-iteration dense LU of the full (2n+m)^2 KKT matrix
// (/root/reference/src/sypha_solver_dense_linear.cpp:150-203: template restore + cusolverDnDgetrf,
// then cusolverDnDgetrs twice) with an m x m Cholesky of M = A D A'.
//
// Layout: row-major, lower triangle, leading dimension ld = m rounded up to 64, identity pad.
//
// Factorisation: k_potrf_df, ONE data-flow launch (left-looking 64x64 tile tasks, see below).  The earlier
// right-looking version (2 launches per 64-wide panel, 500 us at m = 1000) is in the history of this file;
// DESIGN.md 3.2 has the measurements of every step from there to here.
// Solves (k_trsv_df, one launch for forward + backward, data-flow): every 128-row block is a task
//   which accumulates its right-hand side as the blocks it depends on are published through
//   release/acquire flags, then multiplies by the pre-inverted 128x128 diagonal block.  Tasks are
//   claimed in dependency order, so no co-residency is required; spins are bounded and raise an error
//   flag instead of hanging.
#include "sb200_kernels.cuh"
#include "sb200_chol.cuh"
#include "sb200_dmma.cuh"
#include "sb200_tile64.cuh"

#include <cstdlib>

namespace sb200 {

__device__ __forceinline__ void report_fail(int *info, int fail, int base)
{
    if (fail && threadIdx.x == 0)
        atomicCAS(info, 0, base + fail);
}

__global__ void k_pad_identity(int n, double *__restrict__ A, int ld)
{
    // rows n..ld-1: zero, unit diagonal; columns n..ld-1 of the real rows: zero
    const long long total = (long long)ld * ld;
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x)
    {
        const int r = (int)(idx / ld), c = (int)(idx % ld);
        if (r >= n || c >= n)
            A[idx] = (r == c) ? 1.0 : 0.0;
    }
}
void launch_pad_identity(int n, double *a, int ld, cudaStream_t st)
{
    if (ld == n) return;
    k_pad_identity<<<grid_for((long long)ld * ld, 256), 256, 0, st>>>(n, a, ld);
    ++g_launch_count;
}

// ---------------------------------------------------------------------------------------------
// device-side synchronisation shared by the data-flow kernels
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ int ld_acquire(const int *p)
{
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release(int *p, int v)
{
    asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ int ld_volatile(const int *p)
{
    int v;
    asm volatile("ld.volatile.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
// one thread: spin until *flag == epoch.  Bounded: a wait that never completes raises *err (every
// other wait of the launch then falls through) instead of hanging the device.
__device__ __forceinline__ int ld_relaxed(const int *p)
{
    int v;
    asm volatile("ld.relaxed.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
// Polls with relaxed loads (no L1 invalidation per poll) and fences once on success.
__device__ __forceinline__ void spin_until(const int *flag, int epoch, int *err)
{
    int spins = 0;
    if (ld_relaxed(flag) == epoch)
    {
        asm volatile("fence.acq_rel.gpu;" ::: "memory");
        return;
    }
    while (ld_relaxed(flag) != epoch)
    {
        ++spins;
        if ((spins & 255) == 0)
        {
            if (ld_volatile(err)) break;
            if (spins > (1 << 21))
            {
                atomicExch(err, 1);
                break;
            }
        }
    }
    asm volatile("fence.acq_rel.gpu;" ::: "memory");
}
// all threads of the block wrote global data; publish it under `flag`.  The block barrier orders the
// other threads' writes before thread 0's release store (release is cumulative) - no extra fence.
__device__ __forceinline__ void publish(int *flag, int epoch, int publisher = 0)
{
    __syncthreads();
    if (threadIdx.x == publisher) st_release(flag, epoch);
}

#ifdef SB200_DF_TIMING
__device__ unsigned long long g_df_time[8192][12];
__device__ __forceinline__ unsigned long long df_now()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
#define DFT(task, slot) do { if (threadIdx.x == 0 && (task) < 8192) g_df_time[task][slot] = df_now(); } while (0)
#else
#define DFT(task, slot) do { } while (0)
#endif

// A/B switches measured with scripts/df_timeline.cu on B200 (m = 1024, potrf us; baseline 354):
//   SB200_V_RSQ (tile64.cuh) branch-free rsqrt on the pivot chain .............. 321
//   SB200_V_ST   vectorised D1 stores ............................................ 347
//   SB200_V_LJ   L_jj staged through registers (spills when combined) ............ 349
//   SB200_V_PUB  publish of tile (j, j-1) deferred past the diagonal update ...... 357
#ifndef SB200_V_LJ
#define SB200_V_LJ 0
#endif
#ifndef SB200_V_ST
#define SB200_V_ST 1
#endif

// control block of one data-flow kernel family: [0] epoch, [1] next task, [2] exit count
struct DfCtl
{
    int *epoch;
    unsigned *next_task;
    unsigned *exits;
    int *err;
};
// claim the next task (dynamic, in list order: every dependency of a claimed task is owned by a
// running CTA, so no co-residency requirement and no deadlock)
__device__ __forceinline__ int claim_task(const DfCtl &C, int *s_task)
{
    __syncthreads();
    if (threadIdx.x == 0) *s_task = (int)atomicAdd(C.next_task, 1u);
    __syncthreads();
    return *s_task;
}
// last CTA out re-arms the counters and opens the next epoch (=> graph-replayable)
__device__ __forceinline__ void df_epilogue(const DfCtl &C)
{
    __syncthreads();
    if (threadIdx.x == 0)
    {
        __threadfence();
        const unsigned t = atomicAdd(C.exits, 1u);
        if (t == gridDim.x - 1)
        {
            *C.exits = 0u;
            *C.next_task = 0u;
            __threadfence();
            atomicAdd(C.epoch, 1);
        }
    }
}

// ---------------------------------------------------------------------------------------------
// data-flow Cholesky: ONE launch, left-looking per 64x64 tile.
//
// Task list (column-major): chain(j), [pair-inverse((j-1)/2) when j is odd], tile(i,j) for i >= j+2.
// tile(i,j):  acc = A_ij - sum_{k<j} L_ik L_jk'   (waits on the two tiles of column k, in k order);
//             wait D1(j); X = acc L_jj^-T by block substitution with the 16x16 inverses (each warp owns
//             8 rows, no block barrier); write L_ij; publish
// chain(j):   the sub-diagonal tile (j, j-1) AND the diagonal tile (j, j) in one CTA: both accumulate
//             over k < j-1 while waiting; on D1(j-1) the CTA substitutes (j, j-1), publishes it, applies
//             it to the diagonal tile straight from shared memory, factors the tile, inverts its four
//             16x16 diagonal blocks, writes L_jj and those blocks, publishes D1(j); then assembles the
//             full 64x64 inverse (needed by the solves only) and publishes D2
// pair-inverse(b): [W0 0; -W1 L10 W0  W1] = inverse of the 128x128 diagonal block - what the
//             triangular solves multiply by (8 hops at m = 1000 instead of 16).
// The critical path per 64 columns is  factor -> ONE hop -> substitution -> 64^3 DMMA update (no 64x64
// inverse, no kernel boundary, no second hand-off on it).
// ---------------------------------------------------------------------------------------------
struct PotrfDf
{
    double *A;
    int ld, T;
    double *linv;        // [T][64][64]
    double *linv128;     // [(T+1)/2][128][128]
    const int2 *tasks;   // x = i | type << 16, y = j
    int ntasks;
    int *tile_flag;      // [T*T]: tile (i,j) final (diag: D1)
    int *d2_flag;        // [T]
    int *pair_flag;      // [(T+1)/2]
    int *info;
    double2 *d1tag;      // [T][D1_PAIRS]: what the next chain task needs of a factored diagonal tile, tagged
    double *gbuf;        // [T2(T2-1)/2][128][128]: G_ik = W_i L_ik, the blocks the solves stream (see k_trsv_df)
    double *gbufT;       // the same blocks transposed (backward sweep)
    double *zbuf;        // explicit inverse Z = L^-1 (row-major, leading dimension ld) or nullptr: TASK_Z tiles
    double *zTbuf;       // Z'
    int *z_flag;         // [T*T]: Z tile (i,j) final
};
enum { TASK_TILE = 0, TASK_PAIR = 1, TASK_CHAIN = 2, TASK_G = 3, TASK_Z = 4 };
#ifndef SB200_Z_MAX_T
#define SB200_Z_MAX_T 20     // up to 1280 rows the factorisation also forms Z = L^-1 (see TASK_Z)
#endif

__device__ __forceinline__ void ldcg_tile_chunk(double (*S)[KP], const double *g, size_t ld, int tid)
{   // 64 x KC block, bypassing L1 (the tile was written by another SM during this launch)
    for (int idx = tid; idx < TB * (KC / 2); idx += NT_TILE)
    {
        const int r = idx / (KC / 2), c2 = (idx % (KC / 2)) * 2;
        *reinterpret_cast<double2 *>(&S[r][c2]) = __ldcg(reinterpret_cast<const double2 *>(g + (size_t)r * ld + c2));
    }
}

// pair-inverse; operands staged in the two big shared-memory regions (stride XP)
__device__ void pair_inverse_128(unsigned char *smem, const double *__restrict__ Atile10, int ld,
                                 const double *__restrict__ W0g, const double *__restrict__ W1g,
                                 double *__restrict__ out, int tid)
{
    double(*S0)[XP] = reinterpret_cast<double(*)[XP]>(smem + SM_LS);
    double(*S1)[XP] = reinterpret_cast<double(*)[XP]>(smem + SM_LI);
    const int lane = tid & 31, warp = tid >> 5, g = lane >> 2, tg = lane & 3;
    const bool half = (Atile10 == nullptr);    // odd tile count: the pair is [W0 0; 0 I]
    for (int idx = tid; idx < TB * TB / 2; idx += NT_TILE)
    {
        const int r = idx >> 5, c2 = (idx & 31) * 2;
        const double2 w0 = __ldcg(reinterpret_cast<const double2 *>(W0g + r * TB + c2));
        *reinterpret_cast<double2 *>(&S1[r][c2]) = w0;
        *reinterpret_cast<double2 *>(out + (size_t)r * 128 + c2) = w0;
        *reinterpret_cast<double2 *>(out + (size_t)r * 128 + 64 + c2) = make_double2(0.0, 0.0);
        if (!half)
            *reinterpret_cast<double2 *>(&S0[r][c2]) =
                __ldcg(reinterpret_cast<const double2 *>(Atile10 + (size_t)r * ld + c2));
        else
        {
            *reinterpret_cast<double2 *>(out + (size_t)(64 + r) * 128 + c2) = make_double2(0.0, 0.0);
            *reinterpret_cast<double2 *>(out + (size_t)(64 + r) * 128 + 64 + c2) =
                make_double2(r == c2 ? 1.0 : 0.0, r == c2 + 1 ? 1.0 : 0.0);
        }
    }
    if (half) return;
    __syncthreads();
    // Tm = L10 W0 : warp w owns rows 8w..8w+7; W0 is lower triangular => k >= column block start
    double acc[8][2];
#pragma unroll
    for (int nb = 0; nb < 8; ++nb)
    {
        double c0 = 0.0, c1 = 0.0;
        for (int kk = 8 * nb; kk < TB; kk += 4)
            dmma_8x8x4(c0, c1, S0[8 * warp + g][kk + tg], S1[kk + tg][8 * nb + g]);
        acc[nb][0] = c0;
        acc[nb][1] = c1;
    }
    __syncthreads();
#pragma unroll
    for (int nb = 0; nb < 8; ++nb)
        *reinterpret_cast<double2 *>(&S0[8 * warp + g][8 * nb + 2 * tg]) = make_double2(acc[nb][0], acc[nb][1]);
    for (int idx = tid; idx < TB * TB / 2; idx += NT_TILE)
    {
        const int r = idx >> 5, c2 = (idx & 31) * 2;
        const double2 w1 = __ldcg(reinterpret_cast<const double2 *>(W1g + r * TB + c2));
        *reinterpret_cast<double2 *>(&S1[r][c2]) = w1;
        *reinterpret_cast<double2 *>(out + (size_t)(64 + r) * 128 + 64 + c2) = w1;
    }
    __syncthreads();
    // X = -W1 Tm : W1 lower triangular => k <= row block end
#pragma unroll
    for (int nb = 0; nb < 8; ++nb)
    {
        double c0 = 0.0, c1 = 0.0;
        for (int kk = 0; kk < 8 * warp + 8; kk += 4)
            dmma_8x8x4(c0, c1, -S1[8 * warp + g][kk + tg], S0[kk + tg][8 * nb + g]);
        *reinterpret_cast<double2 *>(out + (size_t)(64 + 8 * warp + g) * 128 + 8 * nb + 2 * tg) = make_double2(c0, c1);
    }
}

// G = blockdiag(W_i) L for block row i and tile column tj (both 64x64 tiles of the 128-row block), WITHOUT
// the 128x128 pair inverse: with W_i = [W0 0; -W1 L10 W0, W1],
//     G0 = W0 L(2i, tj)                         needs D2(2i)            (long before the end)
//     T  = L(2i+1, tj) - L10 G0                 needs tile (2i+1, 2i)   (published before chain(2i+1) factors)
//     G1 = W1 T                                 needs D2(2i+1)          (one GEMM after the last inverse)
// so the tail of the factorisation is not lengthened by the solves' operand.  Warp w owns rows 8w..8w+7.
__device__ void g_rows(unsigned char *smem, const PotrfDf &P, const DfCtl &C, int epoch, int i, int tj, int tid)
{
    double(*S0)[XP] = reinterpret_cast<double(*)[XP]>(smem + SM_LS);
    double(*S1)[XP] = reinterpret_cast<double(*)[XP]>(smem + SM_LI);
    const int lane = tid & 31, warp = tid >> 5, g = lane >> 2, tg = lane & 3;
    const int T = P.T, ld = P.ld, t0 = 2 * i, t1 = 2 * i + 1;
    double *out = P.gbuf + ((size_t)i * (i - 1) / 2 + (tj >> 1)) * 128 * 128 + 64 * (tj & 1);
    // transposed copy for the backward sweep (its threads then read contiguous rows, like the forward sweep)
    double *outT = P.gbufT + ((size_t)i * (i - 1) / 2 + (tj >> 1)) * 128 * 128 + (size_t)(64 * (tj & 1)) * 128;
    auto put = [&](const double acc[8][2], int a) {
#pragma unroll
        for (int nb = 0; nb < 8; ++nb)
        {
            *reinterpret_cast<double2 *>(out + (size_t)(64 * a + 8 * warp + g) * 128 + 8 * nb + 2 * tg) =
                make_double2(acc[nb][0], acc[nb][1]);
            outT[(size_t)(8 * nb + 2 * tg) * 128 + 64 * a + 8 * warp + g] = acc[nb][0];
            outT[(size_t)(8 * nb + 2 * tg + 1) * 128 + 64 * a + 8 * warp + g] = acc[nb][1];
        }
    };
    auto stage = [&](double (*S)[XP], const double *src, size_t lds) {
        for (int idx = tid; idx < TB * TB / 2; idx += NT_TILE)
        {
            const int r = idx >> 5, c2 = (idx & 31) * 2;
            *reinterpret_cast<double2 *>(&S[r][c2]) = __ldcg(reinterpret_cast<const double2 *>(src + (size_t)r * lds + c2));
        }
    };
    auto gemm = [&](double acc[8][2], double sign, int kend) {     // acc += sign * S0 S1 (k < kend)
#pragma unroll
        for (int nb = 0; nb < 8; ++nb)
            for (int kk = 0; kk < kend; kk += 4)
                dmma_8x8x4(acc[nb][0], acc[nb][1], sign * S0[8 * warp + g][kk + tg], S1[kk + tg][8 * nb + g]);
    };
    double acc[8][2];
    // ---- G0 = W0 L(2i, tj) ---------------------------------------------------------------------------
    if (tid == 0)
    {
        spin_until(P.d2_flag + t0, epoch, C.err);
        spin_until(P.tile_flag + t0 * T + tj, epoch, C.err);
    }
    __syncthreads();
    stage(S0, P.linv + (size_t)t0 * TB * TB, TB);
    stage(S1, P.A + (size_t)t0 * TB * ld + (size_t)tj * TB, ld);
    __syncthreads();
#pragma unroll
    for (int nb = 0; nb < 8; ++nb)
        acc[nb][0] = acc[nb][1] = 0.0;
    gemm(acc, 1.0, 8 * warp + 8);                      // W0 lower triangular
    put(acc, 0);
    if (t1 >= T) return;                                // odd tile count: the block has one tile row
    __syncthreads();
#pragma unroll
    for (int nb = 0; nb < 8; ++nb)                      // G0 becomes the right operand
        *reinterpret_cast<double2 *>(&S1[8 * warp + g][8 * nb + 2 * tg]) = make_double2(acc[nb][0], acc[nb][1]);
    // ---- T = L(2i+1, tj) - L10 G0 ------------------------------------------------------------------------
    if (tid == 0)
    {
        spin_until(P.tile_flag + t1 * T + t0, epoch, C.err);
        spin_until(P.tile_flag + t1 * T + tj, epoch, C.err);
    }
    __syncthreads();
    stage(S0, P.A + (size_t)t1 * TB * ld + (size_t)t0 * TB, ld);
#pragma unroll
    for (int nb = 0; nb < 8; ++nb)
    {
        const double2 v = __ldcg(reinterpret_cast<const double2 *>(
            P.A + ((size_t)t1 * TB + 8 * warp + g) * ld + (size_t)tj * TB + 8 * nb + 2 * tg));
        acc[nb][0] = v.x;
        acc[nb][1] = v.y;
    }
    __syncthreads();
    gemm(acc, -1.0, TB);
    __syncthreads();
#pragma unroll
    for (int nb = 0; nb < 8; ++nb)
        *reinterpret_cast<double2 *>(&S1[8 * warp + g][8 * nb + 2 * tg]) = make_double2(acc[nb][0], acc[nb][1]);
    // ---- G1 = W1 T -------------------------------------------------------------------------------------
    if (tid == 0) spin_until(P.d2_flag + t1, epoch, C.err);
    __syncthreads();
    stage(S0, P.linv + (size_t)t1 * TB * TB, TB);
    __syncthreads();
#pragma unroll
    for (int nb = 0; nb < 8; ++nb)
        acc[nb][0] = acc[nb][1] = 0.0;
    gemm(acc, 1.0, 8 * warp + 8);
    put(acc, 1);
}

// acc (warp tile 32x16 of the 8-warp 2x4 grid) -= X(rows) X(cols)' with both operands in one
// shared-memory tile of stride XP (full K = 64)
__device__ __forceinline__ void warp_mma_xx(const double (*Xs)[XP], int row0, int col0, int lane, double acc[4][2][2])
{
    const int g = lane >> 2, tg = lane & 3;
#pragma unroll 4
    for (int kk = 0; kk < TB; kk += 4)
    {
        double a[4], b[2];
#pragma unroll
        for (int i = 0; i < 4; ++i)
            a[i] = -Xs[row0 + i * 8 + g][kk + tg];
#pragma unroll
        for (int j = 0; j < 2; ++j)
            b[j] = Xs[col0 + j * 8 + g][kk + tg];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 2; ++j)
                dmma_8x8x4(acc[i][j][0], acc[i][j][1], a[i], b[j]);
    }
}

// Lower-triangle accumulator of a DIAGONAL tile: the 36 8x8 blocks (bi >= bj) of the 64x64 tile are dealt
// to the 8 warps as 5 + 4 per pair of block rows (p, 7 - p) - warp 2p takes the first five of
// [(p,0..p), (7-p,0..7-p)], warp 2p+1 the other four (its fifth slot repeats its last block and is never
// stored).  5 DMMAs per k-step and warp instead of the 8 of the full 32x16 warp tile: the diagonal update
// and the left-looking accumulation of the chain task are bound by the DMMA issue rate.
struct TriMap
{
    int ro[5], co[5];      // first row / first column of slot s
    int nslots;
};
__device__ __forceinline__ TriMap tri_map(int warp)
{
    TriMap m;
    const int p = warp >> 1, h = warp & 1;
    m.nslots = h ? 4 : 5;
#pragma unroll
    for (int s = 0; s < 5; ++s)
    {
        int q = h * 5 + s;                    // index in the pair's list of 9
        if (q > 8) q = 8;
        const int bi = q <= p ? p : 7 - p;
        const int bj = q <= p ? q : q - (p + 1);
        m.ro[s] = 8 * bi;
        m.co[s] = 8 * bj;
    }
    return m;
}
template <int STRIDE>
__device__ __forceinline__ void mma_tri(const double (*S)[STRIDE], int kdepth, const TriMap &m, int lane, double t2[5][2])
{   // t2[s] -= S(rows of slot s) S(cols of slot s)'
    const int g = lane >> 2, tg = lane & 3;
#pragma unroll 2
    for (int kk = 0; kk < kdepth; kk += 4)
    {
        double a[5], b[5];
#pragma unroll
        for (int s = 0; s < 5; ++s)
        {
            a[s] = -S[m.ro[s] + g][kk + tg];
            b[s] = S[m.co[s] + g][kk + tg];
        }
#pragma unroll
        for (int s = 0; s < 5; ++s)
            dmma_8x8x4(t2[s][0], t2[s][1], a[s], b[s]);
    }
}

// X = acc L_jj^-T for tile (ti, tj): block substitution with the 16x16 inverses of L_jj (each warp owns
// 8 rows), result written to the matrix and left in Xs (stride XP).
// TAGGED: the chain's hand-off - L_jj arrives one 16-column PANEL at a time (d1_emit_panel), while chain(tj)
//   is still factoring the panels to the right; step b of the substitution runs as soon as panel b is here, so
//   only the last step (and the last quarter of the diagonal update, `acc2`) follows the end of that factorisation.
//   With `acc2` (the chain's diagonal-tile accumulator) the update acc2 -= X_b X_b' is applied panel by panel
//   for b = 0..2 here; the caller applies panel 3 after publishing X.
// !TAGGED: waits for the D1 flag of tile tj and reads L_jj and the inverses with plain loads.
// All threads must call; ends WITHOUT a barrier (publish() provides it).
__device__ __forceinline__ void trsm_step(double (*Xs)[XP], const double (*Lj)[XP], const double (*Wd)[16][17], int b, int R,
                                          int g, int tg)
{
    double cc[2][2];
#pragma unroll
    for (int nb = 0; nb < 2; ++nb)
    {
        const double2 v = *reinterpret_cast<const double2 *>(&Xs[R + g][16 * b + 8 * nb + 2 * tg]);
        cc[nb][0] = v.x;
        cc[nb][1] = v.y;
    }
    for (int kk = 0; kk < 16 * b; kk += 4)
    {
        const double a = -Xs[R + g][kk + tg];
#pragma unroll
        for (int nb = 0; nb < 2; ++nb)
            dmma_8x8x4(cc[nb][0], cc[nb][1], a, Lj[16 * b + 8 * nb + g][kk + tg]);
    }
    __syncwarp();
#pragma unroll
    for (int nb = 0; nb < 2; ++nb)
        *reinterpret_cast<double2 *>(&Xs[R + g][16 * b + 8 * nb + 2 * tg]) = make_double2(cc[nb][0], cc[nb][1]);
    __syncwarp();
    double o[2][2] = {{0.0, 0.0}, {0.0, 0.0}};
#pragma unroll
    for (int kk = 0; kk < 16; kk += 4)
    {
        const double a = Xs[R + g][16 * b + kk + tg];
        if (kk < 8) dmma_8x8x4(o[0][0], o[0][1], a, Wd[b][g][kk + tg]);     // W lower: k <= n
        dmma_8x8x4(o[1][0], o[1][1], a, Wd[b][8 + g][kk + tg]);
    }
    __syncwarp();
#pragma unroll
    for (int nb = 0; nb < 2; ++nb)
        *reinterpret_cast<double2 *>(&Xs[R + g][16 * b + 8 * nb + 2 * tg]) = make_double2(o[nb][0], o[nb][1]);
    __syncwarp();
}

template <bool TAGGED>
__device__ __forceinline__ void trsm_tile(unsigned char *dyn_smem, const PotrfDf &P, const DfCtl &C, int epoch,
                                          int ti, int tj, const double acc[4][2][2], int t_dbg = 8191,
                                          double (*acc2)[2] = nullptr, const TriMap *tm = nullptr)
{
    double(*Xs)[XP] = reinterpret_cast<double(*)[XP]>(dyn_smem + SM_LS);
    double(*Lj)[XP] = reinterpret_cast<double(*)[XP]>(dyn_smem + SM_LI);
    double(*Wd)[16][17] = reinterpret_cast<double(*)[16][17]>(dyn_smem + SM_T);
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const int row0 = (w >> 2) * 32, col0 = (w & 3) * 16, g = lane >> 2, tg = lane & 3;
    const int T = P.T, ld = P.ld;
    const size_t r0 = (size_t)ti * TB, c0 = (size_t)tj * TB;
    const int R = 8 * w;
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 2; ++j)
            *reinterpret_cast<double2 *>(&Xs[row0 + i * 8 + g][col0 + j * 8 + tg * 2]) =
                make_double2(acc[i][j][0], acc[i][j][1]);
    if (TAGGED)
    {
        const double tag = (double)epoch;
#pragma unroll
        for (int b = 0; b < 4; ++b)
        {   // poll panel b straight into shared memory
            const double2 *src = P.d1tag + (size_t)tj * D1_PAIRS + d1_panel_off(b);
            double2 v[4];
            int spins = 0;
            for (;;)
            {
                bool ok = true;
#pragma unroll
                for (int u = 0; u < 4 - b; ++u)
                {
                    v[u] = ld_tagged(src + tid + u * NT_TILE);
                    ok = ok && (v[u].y == tag);
                }
                if (ok) break;
                if ((++spins & 255) == 0)
                {
                    if (ld_volatile(C.err)) break;
                    if (spins > (1 << 20))
                    {
                        atomicExch(C.err, 1);
                        break;
                    }
                }
            }
#pragma unroll
            for (int u = 0; u < 4 - b; ++u)
            {
                bool is_l;
                int r, c;
                d1_panel_slot(b, tid + u * NT_TILE, is_l, r, c);
                if (is_l)
                    Lj[r][c] = v[u].x;
                else
                    Wd[b][r & 15][c & 15] = v[u].x;
            }
            __syncthreads();                    // panel b staged (b = 0: also the accumulator tile in Xs)
            if (b == 0) DFT(t_dbg, 8);
            if (b == 3) DFT(t_dbg, 9);
            trsm_step(Xs, Lj, Wd, b, R, g, tg);
            if (acc2 && b < 3)
            {   // diagonal update with the 16 columns just finished (every warp's rows are needed)
                __syncthreads();
                mma_tri<XP>(reinterpret_cast<const double(*)[XP]>(&Xs[0][16 * b]), 16, *tm, lane, acc2);
            }
        }
    }
    else
    {
        if (tid == 0) spin_until(P.tile_flag + tj * T + tj, epoch, C.err);
        __syncthreads();
        DFT(t_dbg, 8);
        const double *Ljj = P.A + c0 * ld + c0;
        const double *Wj = P.linv + (size_t)tj * TB * TB;
        for (int idx = tid; idx < TB * TB / 2; idx += NT_TILE)
        {
            const int r = idx >> 5, c2 = (idx & 31) * 2;
            if (c2 <= r)      // the strictly-upper part of L_jj is never read (stale memory there)
                *reinterpret_cast<double2 *>(&Lj[r][c2]) = __ldcg(reinterpret_cast<const double2 *>(Ljj + (size_t)r * ld + c2));
        }
        for (int idx = tid; idx < 4 * 16 * 16; idx += NT_TILE)
        {
            const int b = idx >> 8, r = (idx >> 4) & 15, c = idx & 15;
            Wd[b][r][c] = __ldcg(Wj + (size_t)(16 * b + r) * TB + 16 * b + c);
        }
        __syncthreads();
        DFT(t_dbg, 9);
#pragma unroll
        for (int b = 0; b < 4; ++b)
            trsm_step(Xs, Lj, Wd, b, R, g, tg);
    }
    // each warp writes its own 8 rows (coalesced 512 B rows)
    DFT(t_dbg, 10);
    double *Atile = P.A + r0 * ld + c0;
#pragma unroll
    for (int rr = 0; rr < 8; ++rr)
        *reinterpret_cast<double2 *>(Atile + (size_t)(R + rr) * ld + 2 * lane) =
            *reinterpret_cast<const double2 *>(&Xs[R + rr][2 * lane]);
}

__global__ void __launch_bounds__(NT_TILE, 2) k_potrf_df(PotrfDf P, DfCtl C)
{
    extern __shared__ __align__(16) unsigned char dyn_smem[];
    __shared__ int s_task;
    double(*As)[KP] = reinterpret_cast<double(*)[KP]>(dyn_smem);
    double(*Bs)[KP] = reinterpret_cast<double(*)[KP]>(dyn_smem + TB * KP * 8);
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const int row0 = (w >> 2) * 32, col0 = (w & 3) * 16;     // 8 warps: 2 x 4, warp tile 32 x 16
    const int g = lane >> 2, tg = lane & 3;
    const int T = P.T, ld = P.ld;
    pdl_wait(false);      // no early trigger: CTAs of the successor waiting on the SMs slow the pivot chains (measured)
    const int epoch = ld_volatile(C.epoch) + 1;

    auto load_acc = [&](double acc[4][2][2], size_t r0, size_t c0) {
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 2; ++j)
            {
                const double2 v = __ldcg(reinterpret_cast<const double2 *>(
                    P.A + (r0 + row0 + i * 8 + g) * ld + c0 + col0 + j * 8 + tg * 2));
                acc[i][j][0] = v.x;
                acc[i][j][1] = v.y;
            }
    };

    for (;;)
    {
        const int t = claim_task(C, &s_task);
        if (t >= P.ntasks) break;
        const int2 task = P.tasks[t];
        const int type = task.x >> 16, ti = task.x & 0xffff, tj = task.y;
        DFT(t, 0);

        if (type == TASK_PAIR)
        {
            const int j0 = 2 * tj, j1 = j0 + 1;
            if (tid == 0)
            {
                spin_until(P.d2_flag + j0, epoch, C.err);
                if (j1 < T)
                {
                    spin_until(P.d2_flag + j1, epoch, C.err);
                    spin_until(P.tile_flag + j1 * T + j0, epoch, C.err);
                }
            }
            __syncthreads();
            pair_inverse_128(dyn_smem, j1 < T ? P.A + (size_t)j1 * TB * ld + (size_t)j0 * TB : nullptr, ld,
                             P.linv + (size_t)j0 * TB * TB, P.linv + (size_t)j1 * TB * TB,
                             P.linv128 + (size_t)tj * 128 * 128, tid);
            publish(P.pair_flag + tj, epoch);
            DFT(t, 3);
            continue;
        }

        if (type == TASK_G)
        {   // ---- two 64x64 tiles of G = blockdiag(W) L: block row ti, tile column tj (consumed by the
            //      solves only: visible at kernel end, nothing inside this launch waits for them)
            g_rows(dyn_smem, P, C, epoch, ti, tj, tid);
            DFT(t, 3);
            continue;
        }

        if (type == TASK_Z)
        {   // ---- tile (ti, tj), ti > tj, of Z = L^-1:  Z_ij = -W_i sum_{k=j}^{i-1} L_ik Z_kj  (Z_jj = W_j, written by
            //      chain(j)).  Left-looking like the tiles of L: the sum accumulates in registers as row i of L and
            //      the rows above of column j of Z are published, so that after chain(i) only the product with W_i is
            //      left.  With Z (and Z') in memory the solves M x = b are two triangular matrix-vector products
            //      over the whole GPU, x = Z'(Z b), instead of 2 x T/2 dependent block hops.
            const size_t r0 = (size_t)ti * TB, c0 = (size_t)tj * TB;
            double acc[4][2][2];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int jj = 0; jj < 2; ++jj)
                    acc[i][jj][0] = acc[i][jj][1] = 0.0;
            for (int k = tj; k < ti; ++k)
            {
                if (tid == 0)
                {
                    spin_until(P.tile_flag + ti * T + k, epoch, C.err);
                    spin_until(k == tj ? P.d2_flag + tj : P.z_flag + k * T + tj, epoch, C.err);
                }
                const size_t k0 = (size_t)k * TB;
#pragma unroll
                for (int kc = 0; kc < TB; kc += KC)
                {
                    __syncthreads();
                    ldcg_tile_chunk(As, P.A + r0 * ld + k0 + kc, ld, tid);           // L_ik[:, kc..]
                    ldcg_tile_chunk(Bs, P.zTbuf + c0 * ld + k0 + kc, ld, tid);       // (Z_kj)'[:, kc..]
                    __syncthreads();
                    warp_mma<4, 2>(As, Bs, row0, col0, lane, 1.0, acc);
                }
            }
            __syncthreads();
            DFT(t, 1);
            double(*S0)[XP] = reinterpret_cast<double(*)[XP]>(dyn_smem + SM_LS);
            double(*S1)[XP] = reinterpret_cast<double(*)[XP]>(dyn_smem + SM_LI);
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int jj = 0; jj < 2; ++jj)
                    *reinterpret_cast<double2 *>(&S1[row0 + i * 8 + g][col0 + jj * 8 + tg * 2]) =
                        make_double2(acc[i][jj][0], acc[i][jj][1]);
            if (tid == 0) spin_until(P.d2_flag + ti, epoch, C.err);
            __syncthreads();
            {
                const double *Wi = P.linv + (size_t)ti * TB * TB;
                for (int idx = tid; idx < TB * TB / 2; idx += NT_TILE)
                {
                    const int r = idx >> 5, c2 = (idx & 31) * 2;
                    *reinterpret_cast<double2 *>(&S0[r][c2]) = __ldcg(reinterpret_cast<const double2 *>(Wi + r * TB + c2));
                }
            }
            __syncthreads();
            double *zo = P.zbuf + r0 * ld + c0;            // Z tile (i, j)
            double *zt = P.zTbuf + c0 * ld + r0;           // Z' tile (j, i)
#pragma unroll
            for (int nb = 0; nb < 8; ++nb)
            {   // warp w owns rows 8w..8w+7; W_i lower triangular: k < 8w + 8
                double z0 = 0.0, z1 = 0.0;
                for (int kk = 0; kk < 8 * w + 8; kk += 4)
                    dmma_8x8x4(z0, z1, -S0[8 * w + g][kk + tg], S1[kk + tg][8 * nb + g]);
                *reinterpret_cast<double2 *>(zo + (size_t)(8 * w + g) * ld + 8 * nb + 2 * tg) = make_double2(z0, z1);
                zt[(size_t)(8 * nb + 2 * tg) * ld + 8 * w + g] = z0;
                zt[(size_t)(8 * nb + 2 * tg + 1) * ld + 8 * w + g] = z1;
            }
            publish(P.z_flag + ti * T + tj, epoch);
            DFT(t, 3);
            continue;
        }

        if (type == TASK_TILE)
        {   // ---- regular off-diagonal tile (ti >= tj + 2): left-looking accumulation, then X = acc L_jj^-T
            const size_t r0 = (size_t)ti * TB, c0 = (size_t)tj * TB;
            double acc[4][2][2];
            load_acc(acc, r0, c0);
            for (int k = 0; k < tj; ++k)
            {
                if (tid == 0)
                {
                    spin_until(P.tile_flag + ti * T + k, epoch, C.err);
                    spin_until(P.tile_flag + tj * T + k, epoch, C.err);
                }
                const size_t k0 = (size_t)k * TB;
#pragma unroll
                for (int kc = 0; kc < TB; kc += KC)
                {
                    __syncthreads();
                    ldcg_tile_chunk(As, P.A + r0 * ld + k0 + kc, ld, tid);
                    ldcg_tile_chunk(Bs, P.A + c0 * ld + k0 + kc, ld, tid);
                    __syncthreads();
                    warp_mma<4, 2>(As, Bs, row0, col0, lane, -1.0, acc);
                }
            }
            __syncthreads();
            DFT(t, 1);
            // tile (j+2, j) feeds the chain task two columns on (its last accumulation step waits for it): it takes
            // the tagged hand-off of the diagonal tile like the chain does; the tiles further down have slack and
            // use the flag + plain loads (fewer pollers on the tagged payload)
            if (ti == tj + 2)
                trsm_tile<true>(dyn_smem, P, C, epoch, ti, tj, acc);
            else
                trsm_tile<false>(dyn_smem, P, C, epoch, ti, tj, acc);
            publish(P.tile_flag + ti * T + tj, epoch);
            DFT(t, 3);
            continue;
        }

        // ---- chain task j: sub-diagonal tile (j, j-1) and diagonal tile (j, j) in one CTA, so the only
        //      hand-off on the critical path per 64 columns is D1(j-1) -> here ------------------------
        const int j = tj, jm = tj - 1;
        const size_t c0 = (size_t)j * TB;
        int *deferred = nullptr;
        const TriMap tm = tri_map(w);
        double acc2[5][2];
#pragma unroll
        for (int sl = 0; sl < 5; ++sl)
        {
            const double2 v = __ldcg(reinterpret_cast<const double2 *>(
                P.A + (c0 + tm.ro[sl] + g) * ld + c0 + tm.co[sl] + tg * 2));
            acc2[sl][0] = v.x;
            acc2[sl][1] = v.y;
        }
        if (j > 0)
        {
            const size_t cm = (size_t)jm * TB;
            double acc1[4][2][2];
            load_acc(acc1, c0, cm);
            for (int k = 0; k < jm; ++k)
            {   // the diagonal tile needs only row j of column k; the sub-diagonal tile also row j-1, which for
                // the last k is the tile the previous chain task has just published: it is waited for last, after
                // the diagonal tile's share of this step is done
                const size_t k0 = (size_t)k * TB;
                if (tid == 0) spin_until(P.tile_flag + j * T + k, epoch, C.err);
#pragma unroll
                for (int kc = 0; kc < TB; kc += KC)
                {
                    __syncthreads();
                    ldcg_tile_chunk(kc ? Bs : As, P.A + c0 * ld + k0 + kc, ld, tid);   // row j: both halves staged
                }
                __syncthreads();
                mma_tri<KP>(As, KC, tm, lane, acc2);
                mma_tri<KP>(Bs, KC, tm, lane, acc2);
                if (tid == 0) spin_until(P.tile_flag + jm * T + k, epoch, C.err);
                // acc1 -= L_jk L_(j-1)k' : the halves of L_jk are in As / Bs; stream L_(j-1)k through the scratch tile
                double(*Cs)[KP] = reinterpret_cast<double(*)[KP]>(dyn_smem + SM_LI);
#pragma unroll
                for (int kc = 0; kc < TB; kc += KC)
                {
                    __syncthreads();
                    ldcg_tile_chunk(Cs, P.A + cm * ld + k0 + kc, ld, tid);
                    __syncthreads();
                    warp_mma<4, 2>(kc ? Bs : As, Cs, row0, col0, lane, -1.0, acc1);
                }
            }
            __syncthreads();
            DFT(t, 1);
            // substitution against L_(j-1)(j-1) panel by panel as chain(j-1) emits them, with the diagonal update of
            // panels 0..2 inside; what follows the end of that factorisation is one substitution step and a quarter
            // of the update
            trsm_tile<true>(dyn_smem, P, C, epoch, j, jm, acc1, t, acc2, &tm);
            publish(P.tile_flag + j * T + jm, epoch);           // barrier inside: Xs complete
            DFT(t, 4);
            mma_tri<XP>(reinterpret_cast<const double(*)[XP]>(dyn_smem + SM_LS + 48 * sizeof(double)), 16, tm, lane, acc2);
            __syncthreads();
            DFT(t, 5);
        }
        {   // ---- diagonal tile ----------------------------------------------------------------------
            double(*Ls)[LP] = reinterpret_cast<double(*)[LP]>(dyn_smem + SM_LS);
            double(*Li)[LP] = reinterpret_cast<double(*)[LP]>(dyn_smem + SM_LI);
            // lower blocks only: what is left above the diagonal in Ls (the X tile of the previous step, or
            // uninitialised shared memory for j = 0) is finite-or-not but never reaches a result
#pragma unroll
            for (int sl = 0; sl < 5; ++sl)
                if (sl < tm.nslots)
                {
                    Ls[tm.ro[sl] + g][tm.co[sl] + tg * 2] = acc2[sl][0];
                    Ls[tm.ro[sl] + g][tm.co[sl] + tg * 2 + 1] = acc2[sl][1];
                }
#if defined(SB200_TILE_TIMING) && defined(SB200_TT_TILE)
            if (tid == 0) g_tt_on = (j == SB200_TT_TILE);
#endif
            __syncthreads();
            // chain(j+1) is polling for this tile's panels already: they leave from inside the factorisation
            const int fail = potrf_tile64_factor(dyn_smem, tid, deferred, epoch,
                                                 j + 1 < T ? P.d1tag + (size_t)j * D1_PAIRS : nullptr, (double)epoch);
            DFT(t, 6);
            DFT(t, 7);
            double *Atile = P.A + c0 * ld + c0;
            double *linv_j = P.linv + (size_t)j * TB * TB;
#if SB200_V_ST
#pragma unroll
            for (int u = 0; u < 8; ++u)
            {   // lower triangle of L (the element right of an even diagonal entry is the tile's zero)
                const int idx = tid + u * NT_TILE, r = idx >> 5, c2 = (idx & 31) * 2;
                if (c2 <= r)
                    *reinterpret_cast<double2 *>(Atile + (size_t)r * ld + c2) = make_double2(Ls[r][c2], Ls[r][c2 + 1]);
            }
#pragma unroll
            for (int u = 0; u < 2; ++u)
            {   // the four 16x16 diagonal inverses (the strictly-upper blocks of linv stay zero from allocation)
                const int idx = tid + u * NT_TILE, b = idx >> 7, r = 16 * b + ((idx >> 3) & 15), c2 = 16 * b + (idx & 7) * 2;
                *reinterpret_cast<double2 *>(linv_j + r * TB + c2) = make_double2(Li[r][c2], Li[r][c2 + 1]);
            }
#else
            for (int idx = tid; idx < TB * TB; idx += NT_TILE)
            {
                const int r = idx >> 6, c = idx & 63;
                if (c <= r) Atile[(size_t)r * ld + c] = Ls[r][c];
                if ((r >> 4) == (c >> 4)) linv_j[idx] = Li[r][c];      // diagonal 16-blocks (upper blocks stay zero)
            }
#endif
            report_fail(P.info, fail, (int)c0);
            DFT(t, 11);
            publish(P.tile_flag + j * T + j, epoch);                   // D1
            DFT(t, 2);
            tile64_inv_assemble(dyn_smem, tid);
            for (int idx = tid; idx < TB * TB; idx += NT_TILE)
            {
                const int r = idx >> 6, c = idx & 63;
                if ((r >> 4) > (c >> 4)) linv_j[idx] = Li[r][c];
            }
            if (P.zbuf)
            {   // Z_jj = W_j (zero above the diagonal) and its transpose: the diagonal tiles of Z and Z'
                double *zd = P.zbuf + c0 * ld + c0, *ztd = P.zTbuf + c0 * ld + c0;
                for (int idx = tid; idx < TB * TB; idx += NT_TILE)
                {
                    const int r = idx >> 6, c = idx & 63;
                    zd[(size_t)r * ld + c] = Li[r][c];
                    ztd[(size_t)r * ld + c] = Li[c][r];
                }
            }
            publish(P.d2_flag + j, epoch);                            // D2
            DFT(t, 3);
        }
    }
    df_epilogue(C);
}

// ---------------------------------------------------------------------------------------------
// data-flow triangular solves: ONE launch for both sweeps, 128-row blocks.
//
// With D = blockdiag(L_ii) (128x128 blocks), W = D^-1 and G = W L (unit block diagonal; formed tile by
// tile inside k_potrf_df), M = L L' = D G G' D'.  So M x = b is
//     forward   G y = W b        y_i = (W_i b_i) - sum_{k<i} G_ik y_k
//     backward  G' z = y         z_i = y_i - sum_{k>i} G_ki' z_k
//     scaling   x_i = W_i' z_i
// and a hop of either sweep is ONE 128x128 block times the vector that just arrived, with the block
// already sitting in registers - the two diagonal-inverse products are off the chain (before the
// forward loop, after the backward hand-off).  Tasks 0..T2-1 are the forward blocks (ascending),
// T2..2T2-1 the backward blocks (descending), claimed in dependency order.
// Hand-off without flags or fences: every published value travels as one aligned 16-byte
// {value, epoch tag} store; the consumer polls the pair itself (relaxed 128-bit loads) until the tag
// is this launch's epoch, so a hop costs one store -> L2 -> load round trip.
// ---------------------------------------------------------------------------------------------
static constexpr int NT_TRSV = 512;
static constexpr int WP = 132;                                   // row stride of the inverse block in smem
static constexpr int TRSV_SMEM = 128 * WP * 8;

struct TrsvDf
{
    const double *G, *GT;            // [T2(T2-1)/2][128][128], block (i,k) at i(i-1)/2 + k; GT = blocks transposed
    int ld, T2;
    const double *linv128;
    double *b;
    double2 *fwd_val, *bwd_val;      // [T2*128] tagged y / z
};

__device__ __forceinline__ void trsv_load_w(double (*Ws)[WP], const double *__restrict__ Wg, int tid)
{
    for (int idx = tid; idx < 128 * 64; idx += NT_TRSV)
    {
        const int r = idx >> 6, c2 = (idx & 63) * 2;
        *reinterpret_cast<double2 *>(&Ws[r][c2]) = __ldcg(reinterpret_cast<const double2 *>(Wg + (size_t)r * 128 + c2));
    }
}
// warp 0 polls the 128 tagged values of a block into dst (padded by one per 32), then the block syncs
__device__ __forceinline__ void trsv_fetch(double *dst, const double2 *src, double tag, int *err, int tid)
{
    if (tid < 32)
    {
        double2 v0, v1, v2, v3;
        int spins = 0;
        for (;;)
        {
            v0 = ld_tagged(src + tid);
            v1 = ld_tagged(src + tid + 32);
            v2 = ld_tagged(src + tid + 64);
            v3 = ld_tagged(src + tid + 96);
            if (v0.y == tag && v1.y == tag && v2.y == tag && v3.y == tag) break;
            if ((++spins & 255) == 0)
            {
                if (ld_volatile(err)) break;
                if (spins > (1 << 21))
                {
                    atomicExch(err, 1);
                    break;
                }
            }
        }
        dst[tid] = v0.x;
        dst[tid + 33] = v1.x;
        dst[tid + 66] = v2.x;
        dst[tid + 99] = v3.x;
    }
    __syncthreads();
}

__global__ void __launch_bounds__(NT_TRSV, 1) k_trsv_df(TrsvDf P, DfCtl C)
{
    extern __shared__ __align__(16) unsigned char dyn_smem[];
    __shared__ int s_task;
    __shared__ double ys[2][132];
    __shared__ double yi[132];
    __shared__ double rs[128];
    double(*Ws)[WP] = reinterpret_cast<double(*)[WP]>(dyn_smem);
    const int tid = threadIdx.x, e = tid >> 2, q = tid & 3;
    const int ld = P.ld, T2 = P.T2;
    const int epoch = ld_volatile(C.epoch) + 1;
    const double tag = (double)epoch;

    for (;;)
    {
        const int t = claim_task(C, &s_task);
        if (t >= 2 * T2) break;
        DFT(4096 + t, 0);
        const bool fwd = t < T2;
        const int i = fwd ? t : 2 * T2 - 1 - t;
        const int ge = i * 128 + e;                 // this thread's row (forward) / column (backward)
        trsv_load_w(Ws, P.linv128 + (size_t)i * 128 * 128, tid);
        double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
        double cur[32];
        if (fwd)
        {   // ---- G y = W b: thread (row e, q) owns the column pairs 8j + 2q, 8j + 2q + 1 of its row -------
            const bool rv = ge < ld;
            const double *Grow = P.G + ((size_t)i * (i - 1) / 2) * 128 * 128 + (size_t)e * 128 + 2 * q;
            auto load_cur = [&](int k) {
                const double *src = Grow + (size_t)k * 128 * 128;
#pragma unroll
                for (int j = 0; j < 16; ++j)
                {
                    const double2 v = rv ? __ldcg(reinterpret_cast<const double2 *>(src + 8 * j)) : make_double2(0.0, 0.0);
                    cur[2 * j] = v.x;
                    cur[2 * j + 1] = v.y;
                }
            };
            if (i > 0) load_cur(0);
            if (tid < 128) rs[tid] = (i * 128 + tid < ld) ? __ldcg(P.b + i * 128 + tid) : 0.0;
            __syncthreads();
#pragma unroll
            for (int j = 0; j < 32; j += 4)          // (W_i b_i): columns interleaved over q, conflict-free with WP % 16 == 4
            {
                a0 += Ws[e][4 * j + q] * rs[4 * j + q];
                a1 += Ws[e][4 * j + 4 + q] * rs[4 * j + 4 + q];
                a2 += Ws[e][4 * j + 8 + q] * rs[4 * j + 8 + q];
                a3 += Ws[e][4 * j + 12 + q] * rs[4 * j + 12 + q];
            }
            for (int k = 0; k < i; ++k)
            {
                const double *yk = ys[k & 1];
                trsv_fetch(ys[k & 1], P.fwd_val + (size_t)k * 128, tag, C.err, tid);
#pragma unroll
                for (int j = 0; j < 16; j += 2)
                {   // column c = 8j + 2q (+1) lives at yk[c + (c >> 5)]
                    const int c0 = 8 * j + 2 * q, c1 = 8 * (j + 1) + 2 * q;
                    a0 -= cur[2 * j] * yk[c0 + (c0 >> 5)];
                    a1 -= cur[2 * j + 1] * yk[c0 + 1 + (c0 >> 5)];
                    a2 -= cur[2 * j + 2] * yk[c1 + (c1 >> 5)];
                    a3 -= cur[2 * j + 3] * yk[c1 + 1 + (c1 >> 5)];
                }
                if (k + 1 < i) load_cur(k + 1);
            }
            DFT(4096 + t, 1);
            double yv = (a0 + a1) + (a2 + a3);
            yv += __shfl_xor_sync(0xffffffffu, yv, 1);
            yv += __shfl_xor_sync(0xffffffffu, yv, 2);
            if (q == 0) st_tagged(P.fwd_val + ge, yv, tag);
        }
        else
        {   // ---- G' z = y, then x = W' z: thread (column e, q) owns rows 4j + q of every block below -------
            // row e of the transposed block (k, i) = column e of G_ki; the pair (8j + 2q, 8j + 2q + 1) are rows of
            // block k: both beyond ld (odd tile count) only if the whole pair is (rows come in 64s)
            auto load_cur = [&](int k) {
                const double *src = P.GT + ((size_t)k * (k - 1) / 2 + i) * 128 * 128 + (size_t)e * 128 + 2 * q;
#pragma unroll
                for (int j = 0; j < 16; ++j)
                {
                    const double2 v = (k * 128 + 8 * j + 2 * q < ld) ? __ldcg(reinterpret_cast<const double2 *>(src + 8 * j))
                                                                  : make_double2(0.0, 0.0);
                    cur[2 * j] = v.x;
                    cur[2 * j + 1] = v.y;
                }
            };
            if (i < T2 - 1) load_cur(T2 - 1);
            // y_i (tagged forward result of the same rows) is final long before this block's turn: fetch it
            // now, not on the chain
            trsv_fetch(yi, P.fwd_val + (size_t)i * 128, tag, C.err, tid);
            for (int k = T2 - 1; k > i; --k)
            {
                const double *xk = ys[k & 1];
                trsv_fetch(ys[k & 1], P.bwd_val + (size_t)k * 128, tag, C.err, tid);
                DFT(4096 + t, 4);
#pragma unroll
                for (int j = 0; j < 16; j += 2)
                {
                    const int c0 = 8 * j + 2 * q, c1 = 8 * (j + 1) + 2 * q;
                    a0 -= cur[2 * j] * xk[c0 + (c0 >> 5)];
                    a1 -= cur[2 * j + 1] * xk[c0 + 1 + (c0 >> 5)];
                    a2 -= cur[2 * j + 2] * xk[c1 + (c1 >> 5)];
                    a3 -= cur[2 * j + 3] * xk[c1 + 1 + (c1 >> 5)];
                }
                if (k - 1 > i) load_cur(k - 1);
            }
            DFT(4096 + t, 1);
            double acc = (a0 + a1) + (a2 + a3);
            acc += __shfl_xor_sync(0xffffffffu, acc, 1);
            acc += __shfl_xor_sync(0xffffffffu, acc, 2);
            DFT(4096 + t, 5);
            const double z = yi[e + (e >> 5)] + acc;
            if (q == 0)
            {
                st_tagged(P.bwd_val + ge, z, tag);       // hand-off first: the scaling below is off the chain
                rs[e] = z;
            }
            __syncthreads();
            a0 = a1 = a2 = a3 = 0.0;
#pragma unroll
            for (int j = 0; j < 32; j += 4)
            {
                a0 += Ws[4 * j + q][e] * rs[4 * j + q];
                a1 += Ws[4 * j + 4 + q][e] * rs[4 * j + 4 + q];
                a2 += Ws[4 * j + 8 + q][e] * rs[4 * j + 8 + q];
                a3 += Ws[4 * j + 12 + q][e] * rs[4 * j + 12 + q];
            }
            double xv = (a0 + a1) + (a2 + a3);
            DFT(4096 + t, 6);
            xv += __shfl_xor_sync(0xffffffffu, xv, 1);
            xv += __shfl_xor_sync(0xffffffffu, xv, 2);
            if (q == 0 && ge < ld) P.b[ge] = xv;
        }
        DFT(4096 + t, 3);
    }
    df_epilogue(C);
}

// ---------------------------------------------------------------------------------------------
// triangular matrix-vector product for the explicit-inverse solves: y = Z x over the lower triangle
// (upper = 0: row r spans columns 0..r) or over the upper one (row r spans r..ld-1).  Four rows per CTA, 64
// threads per row, 16-byte loads, every load of a thread independent; fixed summation order (bit-reproducible).
// The element next to the diagonal that a 16-byte access drags in is an explicit zero of the diagonal tile.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_tri_gemv(const double *__restrict__ Z, int ld, const double *__restrict__ x,
                                                  double *__restrict__ y, int upper)
{
    __shared__ double part[8];
    pdl_wait();
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int r = 4 * blockIdx.x + (warp >> 1), t64 = (warp & 1) * 32 + lane;
    const int lo = upper ? (r & ~1) : 0, hi = upper ? ld : ((r + 2) & ~1);
    const double *row = Z + (size_t)r * ld;
    double a0 = 0.0, a1 = 0.0;
#pragma unroll 4
    for (int c = lo + 2 * t64; c < hi; c += 128)
    {
        const double2 z = __ldcg(reinterpret_cast<const double2 *>(row + c));
        a0 = fma(z.x, x[c], a0);          // x: 8-byte loads (caller-owned vector, alignment not assumed)
        a1 = fma(z.y, x[c + 1], a1);
    }
    double a = a0 + a1;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
    if (lane == 0) part[warp] = a;
    __syncthreads();
    if (tid < 4) y[4 * blockIdx.x + tid] = part[2 * tid] + part[2 * tid + 1];
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
static int build_task_list(ErrorSink &err, CholWork &W, int T)
{
    auto hit = W.task_cache.find(T);
    if (hit != W.task_cache.end())
    {
        W.tasks = hit->second.first;
        W.ntasks = hit->second.second;
        W.tasks_T = T;
        return SB200_OK;
    }
    std::vector<int2> tasks;
    auto push_g = [&](int i) {           // G tiles of block row i: one task per tile column below the block
        for (int tj = 0; tj < 2 * i; ++tj)
            tasks.push_back(make_int2(i | (TASK_G << 16), tj));
    };
    int g_done = 0;                      // block rows whose G tasks are listed
    const bool use_z = T <= SB200_Z_MAX_T;
    for (int j = 0; j < T; ++j)
    {
        tasks.push_back(make_int2(j | (TASK_CHAIN << 16), j));      // tile (j, j-1) + diagonal tile j
        if (use_z)
        {   // explicit inverse instead of the pair inverses and G: row j of Z follows column j of L
            for (int i = j + 2; i < T; ++i)
                tasks.push_back(make_int2(i | (TASK_TILE << 16), j));
            for (int c = 0; c < j; ++c)
                tasks.push_back(make_int2(j | (TASK_Z << 16), c));
            continue;
        }
        if ((j & 1) || j == T - 1)
            tasks.push_back(make_int2(TASK_PAIR << 16, j >> 1));    // j even and last: odd tile count
        for (int i = j + 2; i < T; ++i)
            tasks.push_back(make_int2(i | (TASK_TILE << 16), j));
        // block row i is complete after column 2i+1; list its G tasks two columns later: their inputs are
        // final by then, so the CTAs that claim them do not sit on a slot spinning (matters when T is large)
        // (when the task list is longer than the grid can hold at once, they go to the very end instead:
        // interleaved they cost the m = 4096 factorisation 40 %, at the end they are a short tail)
        while (T <= 20 && 2 * g_done + 3 <= j)
            push_g(g_done++);
    }
    while (!use_z && 2 * g_done < T)
        push_g(g_done++);
    W.tasks = nullptr;
    SB200_CUDA_TRY(err, cudaMalloc(&W.tasks, sizeof(int2) * tasks.size()));
    SB200_CUDA_TRY(err, cudaMemcpy(W.tasks, tasks.data(), sizeof(int2) * tasks.size(), cudaMemcpyHostToDevice));
    W.ntasks = (int)tasks.size();
    W.tasks_T = T;
    W.task_cache[T] = std::make_pair(W.tasks, W.ntasks);
    return SB200_OK;
}

int chol_work_ensure(ErrorSink &err, CholWork &W, int n_pad, int n_pad_reserve)
{
    int T = n_pad / TB;
    const int T_now = T, t_reserve = n_pad_reserve / TB;
    if (T > 0xffff) { err.msg = "chol_work_ensure: matrix too large"; return SB200_ERR_UNSUPPORTED; }
    if (T > W.t_cap)
    {
        if (t_reserve > T) T = t_reserve;      // flags / inverse stores sized for the largest model expected
        chol_work_free(W);
        const int T2 = (T + 1) / 2;
        const size_t nflags = 2 * (size_t)T * T + T + 3 * (size_t)T2 + 32;      // ... + Z tile flags at the end
        SB200_CUDA_TRY(err, cudaMalloc(&W.linv, sizeof(double) * (size_t)T * TB * TB));
        SB200_CUDA_TRY(err, cudaMemset(W.linv, 0, sizeof(double) * (size_t)T * TB * TB));
        SB200_CUDA_TRY(err, cudaMalloc(&W.linv128, sizeof(double) * (size_t)T2 * 128 * 128));
        {
            const size_t nblk = (size_t)T2 * (T2 - 1) / 2;
            SB200_CUDA_TRY(err, cudaMalloc(&W.gbuf, sizeof(double) * (nblk ? nblk : 1) * 128 * 128));
            SB200_CUDA_TRY(err, cudaMemset(W.gbuf, 0, sizeof(double) * (nblk ? nblk : 1) * 128 * 128));
            SB200_CUDA_TRY(err, cudaMalloc(&W.gbufT, sizeof(double) * (nblk ? nblk : 1) * 128 * 128));
            SB200_CUDA_TRY(err, cudaMemset(W.gbufT, 0, sizeof(double) * (nblk ? nblk : 1) * 128 * 128));
        }
        {   // Z and Z' for models of up to SB200_Z_MAX_T tiles (the workspace may hold smaller models than T)
            const size_t zl = (size_t)(T < SB200_Z_MAX_T ? T : SB200_Z_MAX_T) * TB;
            SB200_CUDA_TRY(err, cudaMalloc(&W.zbuf, sizeof(double) * zl * zl));
            SB200_CUDA_TRY(err, cudaMemset(W.zbuf, 0, sizeof(double) * zl * zl));
            SB200_CUDA_TRY(err, cudaMalloc(&W.zTbuf, sizeof(double) * zl * zl));
            SB200_CUDA_TRY(err, cudaMemset(W.zTbuf, 0, sizeof(double) * zl * zl));
            SB200_CUDA_TRY(err, cudaMalloc(&W.ytmp, sizeof(double) * zl));
        }
        SB200_CUDA_TRY(err, cudaMalloc(&W.d1tag, sizeof(double2) * (size_t)T * D1_PAIRS));
        SB200_CUDA_TRY(err, cudaMemset(W.d1tag, 0, sizeof(double2) * (size_t)T * D1_PAIRS));
        SB200_CUDA_TRY(err, cudaMalloc(&W.tagged, sizeof(double2) * (size_t)T2 * 128 * 2));
        SB200_CUDA_TRY(err, cudaMemset(W.tagged, 0, sizeof(double2) * (size_t)T2 * 128 * 2));
        SB200_CUDA_TRY(err, cudaMalloc(&W.ctl, sizeof(int) * nflags));
        SB200_CUDA_TRY(err, cudaMemset(W.ctl, 0, sizeof(int) * nflags));
        W.t_cap = T;
        int dev = 0;
        SB200_CUDA_TRY(err, cudaGetDevice(&dev));
        SB200_CUDA_TRY(err, cudaDeviceGetAttribute(&W.sms, cudaDevAttrMultiProcessorCount, dev));
        SB200_CUDA_TRY(err, cudaFuncSetAttribute(k_potrf_df, cudaFuncAttributeMaxDynamicSharedMemorySize, SM_TOTAL));
        SB200_CUDA_TRY(err, cudaFuncSetAttribute(k_trsv_df, cudaFuncAttributeMaxDynamicSharedMemorySize, TRSV_SMEM));
        int occ = 0;
        SB200_CUDA_TRY(err, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_potrf_df, NT_TILE, SM_TOTAL));
        W.potrf_occ = occ < 1 ? 1 : occ;
#ifdef SB200_POTRF_OCC
        W.potrf_occ = SB200_POTRF_OCC;      // A/B: CTAs per SM the data-flow factorisation may use
#endif
    }
    if (T_now != W.tasks_T) return build_task_list(err, W, T_now);
    return SB200_OK;
}
void chol_work_free(CholWork &W)
{
    for (auto &kv : W.task_cache)
        if (kv.second.first) cudaFree(kv.second.first);
    W.task_cache.clear();
    W.tasks = nullptr;
    if (W.linv) cudaFree(W.linv);
    if (W.linv128) cudaFree(W.linv128);
    if (W.ctl) cudaFree(W.ctl);
    if (W.tagged) cudaFree(W.tagged);
    if (W.d1tag) cudaFree(W.d1tag);
    if (W.gbuf) cudaFree(W.gbuf);
    if (W.gbufT) cudaFree(W.gbufT);
    if (W.zbuf) cudaFree(W.zbuf);
    if (W.zTbuf) cudaFree(W.zTbuf);
    if (W.ytmp) cudaFree(W.ytmp);
    W = CholWork{};
}

// ctl layout (ints): [0..2] potrf epoch/next/exits, [4..6] trsv epoch/next/exits, [8] err,
// [32 ..) tile flags T*T, D2 flags T, pair flags T2, fwd flags T2, bwd flags T2   (T = t_cap)
void launch_potrf(CholWork &W, int n, double *a, int ld, int *info, cudaStream_t st)
{
    (void)n;
    const int T = ld / TB, T2 = (T + 1) / 2;
    int *ctl = W.ctl, *flags = W.ctl + 32;
    const size_t tc = (size_t)W.t_cap;
    const bool use_z = T <= SB200_Z_MAX_T;
    PotrfDf P{a, ld, T, W.linv, W.linv128, W.tasks, W.ntasks, flags, flags + tc * tc, flags + tc * tc + tc, info, W.d1tag, W.gbuf, W.gbufT,
              use_z ? W.zbuf : nullptr, use_z ? W.zTbuf : nullptr, flags + tc * tc + tc + 3 * ((tc + 1) / 2)};
    DfCtl C{ctl + 0, reinterpret_cast<unsigned *>(ctl + 1), reinterpret_cast<unsigned *>(ctl + 2), ctl + 8};
    // with the Z tasks in the list (T <= SB200_Z_MAX_T) one CTA per SM is faster than two: a chain task that shares
    // its SM loses issue slots, DMMA cycles and instruction cache to its neighbour (m = 1000: 270 -> 248 us)
    int cap = W.sms * (use_z ? 1 : W.potrf_occ);
    if (W.grid_limit > 0 && W.grid_limit < cap) cap = W.grid_limit;
    const int grid = W.ntasks < cap ? W.ntasks : cap;
    launch_pdl(k_potrf_df, grid, NT_TILE, SM_TOTAL, st, P, C);
    ++g_launch_count;
}

void launch_potrs(CholWork &W, int n, const double *l, int ld, double *b, cudaStream_t st)
{
    (void)n;
    const int T = ld / TB, T2 = (T + 1) / 2;
    int *ctl = W.ctl, *flags = W.ctl + 32;
    const size_t tc = (size_t)W.t_cap, tc2 = (tc + 1) / 2;
    (void)flags;
    (void)l;     // the solves stream G = blockdiag(W) L (or Z = L^-1), built by the factorisation
    if (T <= SB200_Z_MAX_T)
    {   // x = Z'(Z b): two triangular matrix-vector products, four rows per CTA
        launch_pdl(k_tri_gemv, ld / 4, 256, 0, st, W.zbuf, ld, b, W.ytmp, 0);
        launch_pdl(k_tri_gemv, ld / 4, 256, 0, st, W.zTbuf, ld, W.ytmp, b, 1);
        g_launch_count += 2;
        return;
    }
    TrsvDf P{W.gbuf, W.gbufT, ld, T2, W.linv128, b, W.tagged, W.tagged + tc2 * 128};
    DfCtl C{ctl + 4, reinterpret_cast<unsigned *>(ctl + 5), reinterpret_cast<unsigned *>(ctl + 6), ctl + 8};
    const int grid = 2 * T2 < W.sms ? 2 * T2 : W.sms;
    k_trsv_df<<<grid, NT_TRSV, TRSV_SMEM, st>>>(P, C);
    ++g_launch_count;
}

} // namespace sb200
