// sb200_tile64.cuh - Cholesky factor + inverse of one 64x64 tile held in shared memory, one CTA of
// 256 threads.  This routine is the serial spine of the blocked factorisation (it runs once per
// 64-column panel on the critical path), so it is organised around the dependent chain
//     d_c -> rsqrt -> l_cc, l_rc -> d_{c+1}          (measured on B200: rsqrt 75, DFMA 8, SHFL ~25 cycles)
// rather than around throughput:
//   * 4 panels of 16 columns; the 16x16 diagonal block is factored by ONE warp with the rows in
//     registers and warp shuffles (no block barrier inside the 16-column chain);
//   * the rows below are solved one thread per row (right-looking substitution, registers);
//   * the trailing lower triangle is updated by 4x2 register tiles from shared memory;
//   * the inverse of L is assembled afterwards: four 16x16 triangular inverses (one lane per column,
//     no communication), then two levels of inv([A 0; C D]) = [A^-1 0; -D^-1 C A^-1  D^-1].
// 3 block barriers per panel instead of 2 per 4 columns.
#pragma once
#include "sb200_dmma.cuh"

namespace sb200 {

#ifndef SB200_V_LP
#define SB200_V_LP 68
#endif
static constexpr int LP = SB200_V_LP;    // padded row stride of the tile in shared memory (doubles): 65 favours
                                         // one-thread-per-row access, 68 makes the DMMA fragment loads conflict-free

// dynamic shared-memory layout (bytes) of the kernels that factor a diagonal tile
static constexpr int SM_LS = 0;                                   // L tile (aliases the MMA staging buffers)
static constexpr int SM_MMA_BYTES = 2 * TB * KP * 8;              // 36864
static constexpr int SM_LI = SM_MMA_BYTES;                        // inverse tile
static constexpr int XP = TB + 4;                                 // row stride of DMMA-fragment-friendly tiles
static constexpr int SM_LI_BYTES = TB * XP * 8;                   // 34816 (>= TB*LP*8)
static constexpr int SM_T = SM_LI + SM_LI_BYTES;                  // 32x32 scratch of the inverse / 4 x 16x17 blocks
static constexpr int SM_INVD = SM_T + 1088 * 8;                   // 64 reciprocal pivots
static constexpr int SM_FLAG = SM_INVD + 64 * 8;
static constexpr int SM_TOTAL = SM_FLAG + 16;
static constexpr int NT_TILE = 256;                               // threads of the tile factorisation

#ifdef SB200_TILE_TIMING
// stamped by thread 32 (warp 1): a stamp inside warp 0 leaves the pivot-chain warp diverged and makes
// every shuffle take its slow path (measured: 19.6k instead of 3.6k cycles per 16x16 block)
__device__ long long g_tile_timing[64];
__device__ int g_tt_on = 1;          // the chain task switches the stamps on for the tile it wants timed
#define TT(i) do { if (tid == 32 && g_tt_on) g_tile_timing[i] = clock64(); } while (0)
#define TT0(i) do { if (g_tt_on) g_tile_timing[i] = clock64(); } while (0)      // whole warp 0, converged
#else
#define TT(i) do { } while (0)
#define TT0(i) do { } while (0)
#endif

// One 8x8 output block of C = (+/-) A B on the FP64 tensor pipe, operands in shared memory:
// A row-major [i][k] (lda), B row-major [k][j] (ldb), K a multiple of 4.  Executed by one full warp.
template <bool NEGATE>
__device__ __forceinline__ void dmma_block_nn(const double *A, int lda, const double *B, int ldb, double *C,
                                              int ldc, int K, int lane)
{
    const int g = lane >> 2, tg = lane & 3;
    double c0 = 0.0, c1 = 0.0;
    for (int kk = 0; kk < K; kk += 4)
        dmma_8x8x4(c0, c1, A[g * lda + kk + tg], B[(kk + tg) * ldb + g]);
    C[g * ldc + 2 * tg] = NEGATE ? -c0 : c0;
    C[g * ldc + 2 * tg + 1] = NEGATE ? -c1 : c1;
}

// C (8x8 block, in place) -= A A2' with both operands row-major [row][k]: the Cholesky trailing update
__device__ __forceinline__ void dmma_block_nt_sub(const double *A, const double *A2, int lda, double *C, int ldc,
                                                  int K, int lane)
{
    const int g = lane >> 2, tg = lane & 3;
    double c0 = C[g * ldc + 2 * tg], c1 = C[g * ldc + 2 * tg + 1];
    for (int kk = 0; kk < K; kk += 4)
        dmma_8x8x4(c0, c1, -A[g * lda + kk + tg], A2[g * lda + kk + tg]);
    C[g * ldc + 2 * tg] = c0;
    C[g * ldc + 2 * tg + 1] = c1;
}

// 1/sqrt(x) for a normal positive double: MUFU seed (about 20 bits) + one third-order correction
// y0 (1 + e/2 + 3e^2/8), e = 1 - x y0^2 - the same arithmetic as CUDA's rsqrt() without its
// special-case branch, which costs more than the arithmetic on the serial pivot chain (isolated
// 16x16 block on B200: 3630 -> 2420 cycles).  Non-positive / non-finite pivots are caught by the
// caller's d > 0 test; the garbage this returns for them never reaches a reported result.
__device__ __forceinline__ double rsqrt_pivot(double x)
{
    double y0;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y0) : "d"(x));
    const double e = fma(-(y0 * y0), x, 1.0);
    return fma(fma(e, 0.375, 0.5), y0 * e, y0);
}

#ifndef SB200_V_RSQ
#define SB200_V_RSQ 1
#endif
#ifndef SB200_V_TRAIL
#define SB200_V_TRAIL 1      // three interleaved blocks per warp (straight-line form; the first form with per-slot
                             // branches and pointer arrays measured slower: 354 -> 395 us at m = 1024)
#endif
#if SB200_V_RSQ
#define SB200_RSQ(x) rsqrt_pivot(x)
#else
#define SB200_RSQ(x) rsqrt(x)
#endif

// Factor: returns 0 or the 1-based local index of the first non-positive pivot (same value in all
// threads).  On exit Ls = L (zero above the diagonal inside the 16x16 diagonal blocks, unspecified in the
// strictly-upper 16x16 blocks) and the four 16x16 diagonal blocks of Li hold the inverses
// of the diagonal blocks of L (the strictly-upper 16x16 blocks of Li are cleared).
//
// Per 16-column panel:
//   1. ONE warp factors the 16x16 diagonal block: lanes 0..15 hold its rows in registers, column c of
//      L is broadcast through shared memory, and the pivot chain dg -> shfl -> rsqrt -> l is issued one
//      column ahead of the bulk update (measured on B200: shfl 26, rsqrt 67, DFMA 8 cycles).  Lanes
//      16..31 run the SAME instruction stream on the columns of the identity, which yields W = L_dd^-1
//      at no extra latency.
//   2. the rows below become X = A W' on the tensor pipe (one 8-row block per warp) instead of a
//      16-step substitution;
//   3. the trailing lower triangle is updated with unrolled 8x8x16 DMMA blocks.
__device__ __forceinline__ void st_release_gpu(int *p, int v)
{
    asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// {value, epoch tag} pairs: one aligned 16-byte transaction, polled by the consumer itself (no flag, no
// fence): 457 ns per hand-off against 978 ns for data + release flag + acquire fence (scripts/pingpong.cu)
__device__ __forceinline__ void st_tagged(double2 *p, double v, double tag)
{
    asm volatile("st.relaxed.gpu.global.v2.f64 [%0], {%1, %2};" ::"l"(p), "d"(v), "d"(tag) : "memory");
}
__device__ __forceinline__ double2 ld_tagged(const double2 *p)
{
    double2 v;
    asm volatile("ld.relaxed.gpu.global.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p) : "memory");
    return v;
}

// What the next chain task needs of a factored diagonal tile travels as tagged pairs, one PANEL (16 columns) at
// a time, as soon as the panel is final: panel b = the 16x16 blocks (b+1..3, b) of L_jj, then W_b = L_bb^-1.
// The consumer substitutes against panel b while the producer is still factoring panels b+1..3.
static constexpr int D1_PAIRS = 6 * 256 + 4 * 256;
__device__ __forceinline__ int d1_panel_off(int b) { return b == 0 ? 0 : (b == 1 ? 1024 : (b == 2 ? 1792 : 2304)); }
__device__ __forceinline__ void d1_panel_slot(int b, int q, bool &is_l, int &r, int &c)
{
    const int blk = q >> 8, e = q & 255;
    is_l = blk < 3 - b;
    r = (is_l ? 16 * (b + 1 + blk) : 16 * b) + (e >> 4);
    c = 16 * b + (e & 15);
}
__device__ __forceinline__ void d1_emit_panel(double2 *dst, double tag, int b, const double (*Ls)[LP], const double (*Li)[LP],
                                              int t, int nthreads)
{
    double2 *out = dst + d1_panel_off(b);
    for (int q = t; q < (4 - b) * 256; q += nthreads)
    {
        bool is_l;
        int r, c;
        d1_panel_slot(b, q, is_l, r, c);
        st_tagged(out + q, is_l ? Ls[r][c] : Li[r][c], tag);
    }
}

#ifndef SB200_V_IDLE_WARP4
#define SB200_V_IDLE_WARP4 1
#endif
#ifndef SB200_V_LOOKAHEAD
#define SB200_V_LOOKAHEAD 1  // the pivot-chain warp also applies a finished 16-column panel to the NEXT 16x16 diagonal
                             // block (all its next chain needs); the other seven warps apply it to the rest of the
                             // tile one step later, beside that chain
#endif

__device__ __forceinline__ void named_bar_sync(int id, int count)
{
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory");
}
__device__ __forceinline__ void named_bar_arrive(int id, int count)
{
    asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(count) : "memory");
}

// The 16x16 diagonal block at c0 of Ls factored by ONE warp (see potrf_tile64_factor below): lanes 0..15 hold
// the rows, lanes 16..31 the columns of the identity (-> W = L_dd^-1 in Li).  Returns the 1-based tile-local
// index of the first non-positive pivot, or 0.
#ifndef SB200_V_P_NOINLINE
#define SB200_V_P_NOINLINE 1   // one copy of the 16-column chain in the kernel: inlined, the compiler peels the panel loop
                               // and the second copy is a second set of cold instruction-cache lines per tile
#endif
#if SB200_V_P_NOINLINE
__device__ __noinline__ int potrf_block16_warp(double (*Ls)[LP], double (*Li)[LP], double *Tb, int c0, int lane)
#else
__device__ __forceinline__ int potrf_block16_warp(double (*Ls)[LP], double (*Li)[LP], double *Tb, int c0, int lane)
#endif
{
    const int r = lane & 15;
    const bool inv_lane = lane >= 16;
    double(*Cb)[17] = reinterpret_cast<double(*)[17]>(Tb);
    double a[16];
    {   // rows of the block, or (inverse lanes) rows of the 16x16 identity kept behind the column buffer: one
        // pointer select and eight 16-byte loads instead of a compare-and-select per entry
        const double *src = inv_lane ? Tb + 512 + 16 * r : &Ls[c0 + r][c0];
#pragma unroll
        for (int j = 0; j < 16; j += 2)
        {
            const double2 v = *reinterpret_cast<const double2 *>(src + j);
            a[j] = v.x;
            a[j + 1] = v.y;
        }
    }
    double dg = Ls[c0 + r][c0 + r];            // this lane's own diagonal entry (one load instead of 15 selects; the
                                               // inverse lanes never contribute theirs)
    unsigned badmask = 0;                      // bit c: pivot c not positive (the index is worked out after the chain)
    double d = __shfl_sync(0xffffffffu, dg, 0);
    double inv = SB200_RSQ(d);
#pragma unroll
    for (int c = 0; c < 16; ++c)
    {
        if (!(d > 0.0)) badmask |= 1u << c;
        // rows: L[r][c] (for r == c the row's own a[c] has received exactly the updates of dg, so a[c] * inv is
        // d * inv = sqrt(d) without a special case);  inverse lanes: z_c = z[c] / L[c][c]
        const double l = a[c] * inv;
        a[c] = l;
        dg -= l * l;
        if (!inv_lane) Cb[c][r] = l;
        if (c < 15)
        {
            d = __shfl_sync(0xffffffffu, dg, c + 1);
            inv = SB200_RSQ(d);
        }
        __syncwarp();
#pragma unroll
        for (int c2 = c + 1; c2 < 16; ++c2)
            a[c2] -= l * Cb[c][c2];
    }
    if (!inv_lane)
    {
#pragma unroll
        for (int j = 0; j < 16; ++j)
            Ls[c0 + r][c0 + j] = (j <= r) ? a[j] : 0.0;
    }
    else
    {
#pragma unroll
        for (int j = 0; j < 16; ++j)
            Li[c0 + j][c0 + r] = a[j];
    }
    return badmask ? c0 + __ffs(badmask) : 0;
}

#if SB200_V_LOOKAHEAD
// Look-ahead form of the tile factorisation.  With 16x16 blocks A[i][j] (i >= j) of the tile and panel k =
// block column k, the dependent chain is
//     P(k): chol(A[k][k]) -> N(k): L[k+1][k] = A[k+1][k] W_k', A[k+1][k+1] -= L[k+1][k] L[k+1][k]' -> P(k+1)
// and everything else panel k has to do - R(k): L[i][k] = A[i][k] W_k' for i >= k+2, T(k): A[i][j] -= L[i][k]
// L[j][k]' for i >= j >= k+1 except (k+1, k+1) - is only needed by N(k+1).  Warp 0 runs P and N back to back;
// warps 1..7 run R(k-1) and T(k-1) during P(k) (step k), so a panel costs one pivot chain plus two 16x16x16
// products instead of a chain, a rows-below phase and a trailing phase separated by block barriers.  Every
// element still receives its panel updates in panel order with the same DMMA sequences as the phased form.
//   barriers per step: named 1 (warps 1..7, between R and T), named 2 (warps 1..7 arrive after T, warp 0 waits
//   before N), block barrier at the end of the step (W_k, L[k+1][k] visible to warps 1..7).
// `d1dst` (may be null): warps 1..7 send panel k-1 of the tagged hand-off (d1_emit_panel) in step k, once their
// share of the step is done - the next chain task substitutes against it while this tile is still being factored.
__device__ int potrf_tile64_factor(unsigned char *smem, int tid, int *deferred_flag = nullptr, int deferred_value = 0,
                                   double2 *d1dst = nullptr, double d1tag = 0.0)
{
    TT(0);
    double(*Ls)[LP] = reinterpret_cast<double(*)[LP]>(smem + SM_LS);
    double(*Li)[LP] = reinterpret_cast<double(*)[LP]>(smem + SM_LI);
    double *Tb = reinterpret_cast<double *>(smem + SM_T);
    int *sflag = reinterpret_cast<int *>(smem + SM_FLAG);
    const int lane = tid & 31, warp = tid >> 5, g = lane >> 2, tg = lane & 3;
    if (tid == 0) *sflag = 0;
#ifdef SB200_SKIP_FACTOR
    __syncthreads();
    return 0;
#endif
    Tb[512 + tid] = ((tid >> 4) == (tid & 15)) ? 1.0 : 0.0;      // 16x16 identity for the inverse lanes (NT_TILE = 256)
    for (int idx = tid; idx < 6 * 256; idx += NT_TILE)
    {   // clear the strictly-upper 16x16 blocks of Li: (0,1) (0,2) (0,3) (1,2) (1,3) (2,3)
        const int blk = idx >> 8, e = idx & 255;
        const int bi = blk < 3 ? 0 : (blk < 5 ? 1 : 2);
        const int bj = blk < 3 ? blk + 1 : (blk < 5 ? blk - 1 : 3);
        Li[16 * bi + (e >> 4)][16 * bj + (e & 15)] = 0.0;
    }
    __syncthreads();
    if (tid == 32 && deferred_flag) st_release_gpu(deferred_flag, deferred_value);

#pragma unroll 1
    for (int k = 0; k < 4; ++k)
    {
        const int c0 = 16 * k;
        TT(8 + 4 * k);
        if (warp == 0)
        {
            // ---- P(k) ------------------------------------------------------------------------------------
            const int bad = potrf_block16_warp(Ls, Li, Tb, c0, lane);
            if (lane == 0 && bad) atomicCAS(sflag, 0, bad);
            __syncwarp();
            TT0(10 + 4 * k);
            if (k < 3)
            {
                named_bar_sync(2, NT_TILE);                 // T(k-1) done: A[k+1][k] and A[k+1][k+1] are current
                // ---- N(k): X = A[k+1][k] W_k' (W lower triangular: output columns 0..7 need k < 8 only) ----
                const int R = c0 + 16;
                double af[2][4];
#pragma unroll
                for (int rb = 0; rb < 2; ++rb)
#pragma unroll
                    for (int kk = 0; kk < 4; ++kk)
                        af[rb][kk] = Ls[R + 8 * rb + g][c0 + 4 * kk + tg];
                double x[2][2][2] = {};
#pragma unroll
                for (int kk = 0; kk < 4; ++kk)
                {
                    const double w1 = Li[c0 + 8 + g][c0 + 4 * kk + tg];
                    if (kk < 2)
                    {
                        const double w0 = Li[c0 + g][c0 + 4 * kk + tg];
                        dmma_8x8x4(x[0][0][0], x[0][0][1], af[0][kk], w0);
                        dmma_8x8x4(x[1][0][0], x[1][0][1], af[1][kk], w0);
                    }
                    dmma_8x8x4(x[0][1][0], x[0][1][1], af[0][kk], w1);
                    dmma_8x8x4(x[1][1][0], x[1][1][1], af[1][kk], w1);
                }
                __syncwarp();
#pragma unroll
                for (int rb = 0; rb < 2; ++rb)
#pragma unroll
                    for (int cb = 0; cb < 2; ++cb)
                    {
                        Ls[R + 8 * rb + g][c0 + 8 * cb + 2 * tg] = x[rb][cb][0];
                        Ls[R + 8 * rb + g][c0 + 8 * cb + 2 * tg + 1] = x[rb][cb][1];
                    }
                __syncwarp();
                // A[k+1][k+1] -= X X' (lower 8x8 blocks (0,0), (1,0), (1,1))
                double xa[2][4], xb[2][4];
#pragma unroll
                for (int rb = 0; rb < 2; ++rb)
#pragma unroll
                    for (int kk = 0; kk < 4; ++kk)
                    {
                        xb[rb][kk] = Ls[R + 8 * rb + g][c0 + 4 * kk + tg];
                        xa[rb][kk] = -xb[rb][kk];
                    }
                double u00[2], u10[2], u11[2];
                u00[0] = Ls[R + g][R + 2 * tg];          u00[1] = Ls[R + g][R + 2 * tg + 1];
                u10[0] = Ls[R + 8 + g][R + 2 * tg];      u10[1] = Ls[R + 8 + g][R + 2 * tg + 1];
                u11[0] = Ls[R + 8 + g][R + 8 + 2 * tg];  u11[1] = Ls[R + 8 + g][R + 8 + 2 * tg + 1];
#pragma unroll
                for (int kk = 0; kk < 4; ++kk)
                {
                    dmma_8x8x4(u00[0], u00[1], xa[0][kk], xb[0][kk]);
                    dmma_8x8x4(u10[0], u10[1], xa[1][kk], xb[0][kk]);
                    dmma_8x8x4(u11[0], u11[1], xa[1][kk], xb[1][kk]);
                }
                __syncwarp();
                Ls[R + g][R + 2 * tg] = u00[0];          Ls[R + g][R + 2 * tg + 1] = u00[1];
                Ls[R + 8 + g][R + 2 * tg] = u10[0];      Ls[R + 8 + g][R + 2 * tg + 1] = u10[1];
                Ls[R + 8 + g][R + 8 + 2 * tg] = u11[0];  Ls[R + 8 + g][R + 8 + 2 * tg + 1] = u11[1];
                TT0(11 + 4 * k);
            }
        }
        else
        {
#if SB200_V_IDLE_WARP4
            // warp 4 shares its scheduler (SM sub-partition 0) with the pivot-chain warp: it only keeps the barrier
            // counts and leaves the issue slots to warp 0; the other six warps split the work
            const bool worker = warp != 4;
            const int w = warp - 1 - (warp > 4 ? 1 : 0);    // 0..5 for the workers
            constexpr int NWORK = 6;
#else
            const bool worker = true;
            const int w = warp - 1;                         // 0..6
            constexpr int NWORK = 7;
#endif
            if (k >= 1 && k <= 2)
            {
                const int p0 = c0 - 16;                     // panel k-1
                // ---- R(k-1): rows p0+32..63, one 8-row block per warp ------------------------------------------
                const int nrb = (32 - p0) >> 3;             // 4, 2
                if (worker && w < nrb)
                {
                    const int R = p0 + 32 + 8 * w;
                    double af[4];
#pragma unroll
                    for (int kk = 0; kk < 4; ++kk)
                        af[kk] = Ls[R + g][p0 + 4 * kk + tg];
                    double x00 = 0.0, x01 = 0.0, x10 = 0.0, x11 = 0.0;
#pragma unroll
                    for (int kk = 0; kk < 4; ++kk)
                    {
                        if (kk < 2) dmma_8x8x4(x00, x01, af[kk], Li[p0 + g][p0 + 4 * kk + tg]);
                        dmma_8x8x4(x10, x11, af[kk], Li[p0 + 8 + g][p0 + 4 * kk + tg]);
                    }
                    __syncwarp();
                    Ls[R + g][p0 + 2 * tg] = x00;
                    Ls[R + g][p0 + 2 * tg + 1] = x01;
                    Ls[R + g][p0 + 8 + 2 * tg] = x10;
                    Ls[R + g][p0 + 8 + 2 * tg + 1] = x11;
                }
                named_bar_sync(1, NT_TILE - 32);
                if (worker)
                {
                // ---- T(k-1): 8x8 blocks (bi, bj), bi >= bj, of rows/cols r0..63 except the first 16x16 block ---
                    const int r0 = p0 + 16, nb = (64 - r0) >> 3;     // 6, 4
                    const int nblk = nb * (nb + 1) / 2 - 3;          // 18, 7
                    int ro[3], co[3];
    #pragma unroll
                    for (int v = 0; v < 3; ++v)
                    {
                        const int q = w + NWORK * v + 3;             // skip (0,0) (1,0) (1,1)
                        int bi = 0;
                        bi += (q >= 1) + (q >= 3) + (q >= 6) + (q >= 10) + (q >= 15);
                        const int bj = q - bi * (bi + 1) / 2;
                        const bool act = worker && w + NWORK * v < nblk;
                        ro[v] = act ? r0 + 8 * bi : r0 + 16;
                        co[v] = act ? r0 + 8 * bj : r0;
                    }
                    double af[3][4], bf[3][4], u[3][2];
    #pragma unroll
                    for (int v = 0; v < 3; ++v)
                    {
    #pragma unroll
                        for (int kk = 0; kk < 4; ++kk)
                        {
                            af[v][kk] = -Ls[ro[v] + g][p0 + 4 * kk + tg];
                            bf[v][kk] = Ls[co[v] + g][p0 + 4 * kk + tg];
                        }
                        u[v][0] = Ls[ro[v] + g][co[v] + 2 * tg];
                        u[v][1] = Ls[ro[v] + g][co[v] + 2 * tg + 1];
                    }
    #pragma unroll
                    for (int kk = 0; kk < 4; ++kk)
    #pragma unroll
                        for (int v = 0; v < 3; ++v)
                            dmma_8x8x4(u[v][0], u[v][1], af[v][kk], bf[v][kk]);
                    __syncwarp();
    #pragma unroll
                    for (int v = 0; v < 3; ++v)
                        if (worker && w + NWORK * v < nblk)
                        {
                            Ls[ro[v] + g][co[v] + 2 * tg] = u[v][0];
                            Ls[ro[v] + g][co[v] + 2 * tg + 1] = u[v][1];
                        }
                }
            }
            TT(9 + 4 * k);
            if (k < 3)
            {
                __threadfence_block();
                named_bar_arrive(2, NT_TILE);
            }
            if (d1dst && k >= 1 && worker) d1_emit_panel(d1dst, d1tag, k - 1, Ls, Li, 32 * w + lane, 32 * NWORK);
            TT(24 + k);
        }
        __syncthreads();
    }
    if (d1dst) d1_emit_panel(d1dst, d1tag, 3, Ls, Li, tid, NT_TILE);
    TT(1);
    return *sflag;
}
#else
// `deferred_flag`: a publish the caller still owes (its data was stored and a block barrier has passed): a
// thread of warp 1 releases it while warp 0 runs the first pivot chain, so the release fence (~1 us) is
// hidden instead of delaying the caller's critical path.
__device__ int potrf_tile64_factor(unsigned char *smem, int tid, int *deferred_flag = nullptr, int deferred_value = 0,
                                   double2 *d1dst = nullptr, double d1tag = 0.0)
{
    TT(0);
    double(*Ls)[LP] = reinterpret_cast<double(*)[LP]>(smem + SM_LS);
    double(*Li)[LP] = reinterpret_cast<double(*)[LP]>(smem + SM_LI);
    double *Tb = reinterpret_cast<double *>(smem + SM_T);        // the 16x16 column buffer of the diagonal block
    int *sflag = reinterpret_cast<int *>(smem + SM_FLAG);
    const int lane = tid & 31, warp = tid >> 5, g = lane >> 2, tg = lane & 3;
    if (tid == 0) *sflag = 0;
#ifdef SB200_SKIP_FACTOR
    __syncthreads();
    return 0;
#endif
    Tb[512 + tid] = ((tid >> 4) == (tid & 15)) ? 1.0 : 0.0;      // 16x16 identity for the inverse lanes (NT_TILE = 256)
    // clear the strictly-upper 16x16 blocks of Li: (0,1) (0,2) (0,3) (1,2) (1,3) (2,3)
    for (int idx = tid; idx < 6 * 256; idx += NT_TILE)
    {
        const int blk = idx >> 8, e = idx & 255;
        const int bi = blk < 3 ? 0 : (blk < 5 ? 1 : 2);
        const int bj = blk < 3 ? blk + 1 : (blk < 5 ? blk - 1 : 3);
        Li[16 * bi + (e >> 4)][16 * bj + (e & 15)] = 0.0;
    }

#pragma unroll 1
    for (int kb = 0; kb < 4; ++kb)
    {
        const int c0 = 16 * kb;
        __syncthreads();
        TT(8 + 4 * kb);
        if (kb == 0 && tid == 32 && deferred_flag) st_release_gpu(deferred_flag, deferred_value);
#ifndef SB200_SKIP_DIAG
        if (warp == 0)
#else
        if (false)
#endif
        {
            const int r = lane & 15;
            const bool inv_lane = lane >= 16;          // lanes 16..31: column r of W = L_dd^-1
            double(*Cb)[17] = reinterpret_cast<double(*)[17]>(Tb);
            double a[16];
#pragma unroll
            for (int j = 0; j < 16; ++j)
                a[j] = inv_lane ? (j == r ? 1.0 : 0.0) : Ls[c0 + r][c0 + j];
            double dg = a[0];
#pragma unroll
            for (int j = 1; j < 16; ++j)
                dg = (r == j) ? a[j] : dg;             // this lane's own diagonal entry
            int bad = 0;
            double d = __shfl_sync(0xffffffffu, dg, 0);
            double inv = SB200_RSQ(d);
#pragma unroll
            for (int c = 0; c < 16; ++c)
            {
                if (!(d > 0.0) && bad == 0) bad = c0 + c + 1;
                // rows: L[r][c] (meaningful for r >= c);  inverse lanes: z_c = z[c] / L[c][c]
                const double l = (!inv_lane && r == c) ? d * inv : a[c] * inv;
                a[c] = l;
                dg -= l * l;
                if (!inv_lane) Cb[c][r] = l;
                if (c < 15)
                {   // next pivot: long-latency chain started before this column's bulk update
                    d = __shfl_sync(0xffffffffu, dg, c + 1);
                    inv = SB200_RSQ(d);
                }
                __syncwarp();
#pragma unroll
                for (int c2 = c + 1; c2 < 16; ++c2)
                    a[c2] -= l * Cb[c][c2];            // L[c2][c] (broadcast read); inverse lanes: z[c2] -= L[c2][c] z_c
            }
            if (!inv_lane)
            {
#pragma unroll
                for (int j = 0; j < 16; ++j)
                    Ls[c0 + r][c0 + j] = (j <= r) ? a[j] : 0.0;
            }
            else
            {
#pragma unroll
                for (int j = 0; j < 16; ++j)
                    Li[c0 + j][c0 + r] = a[j];         // W[j][r]; zero above the diagonal by construction
            }
            if (lane == 0 && bad) atomicCAS(sflag, 0, bad);
        }
        __syncthreads();
        TT(9 + 4 * kb);
        const int nrows = 48 - c0;
#ifdef SB200_SKIP_ROWS
        if (false)
#else
        if (8 * warp < nrows)
#endif
        {   // ---- rows below: X = A W' (W lower triangular: k <= n), one 8-row block per warp ----------
            const int R = c0 + 16 + 8 * warp;
            double af[4];
#pragma unroll
            for (int kk = 0; kk < 4; ++kk)
                af[kk] = Ls[R + g][c0 + 4 * kk + tg];
            double x00 = 0.0, x01 = 0.0, x10 = 0.0, x11 = 0.0;
#pragma unroll
            for (int kk = 0; kk < 4; ++kk)
            {
                if (kk < 2) dmma_8x8x4(x00, x01, af[kk], Li[c0 + g][c0 + 4 * kk + tg]);
                dmma_8x8x4(x10, x11, af[kk], Li[c0 + 8 + g][c0 + 4 * kk + tg]);
            }
            __syncwarp();
            Ls[R + g][c0 + 2 * tg] = x00;
            Ls[R + g][c0 + 2 * tg + 1] = x01;
            Ls[R + g][c0 + 8 + 2 * tg] = x10;
            Ls[R + g][c0 + 8 + 2 * tg + 1] = x11;
        }
        // (the part of Ls right of the diagonal blocks is never read: it used to be zeroed here by one warp with
        //  a runtime division per element - 1.5 us per tile on the critical path, found by skipping phases)
        __syncthreads();
        TT(10 + 4 * kb);
#ifdef SB200_SKIP_TRAIL
        if (false)
#endif
#if SB200_V_TRAIL
        {   // ---- trailing update on the tensor pipe: 8x8 blocks of the lower triangle; a warp owns up to
            //      three blocks (q = warp, warp+8, warp+16).  Straight-line code: all fragment loads first,
            //      then the three independent DMMA chains interleaved; a slot without a block works on
            //      block (0,0) and simply does not store.
            const int r0 = c0 + 16, nb = nrows >> 3, nblk = nb * (nb + 1) / 2;
            if (nblk > 0)
            {
                int ro[3], co[3];
#pragma unroll
                for (int v = 0; v < 3; ++v)
                {
                    const int q = warp + 8 * v;
                    int bi = 0;
                    bi += (q >= 1) + (q >= 3) + (q >= 6) + (q >= 10) + (q >= 15);
                    const int bj = q - bi * (bi + 1) / 2;
                    const bool act = q < nblk;
                    ro[v] = act ? r0 + 8 * bi : r0;
                    co[v] = act ? r0 + 8 * bj : r0;
                }
                double af[3][4], bf[3][4], u[3][2];
#pragma unroll
                for (int v = 0; v < 3; ++v)
                {
#pragma unroll
                    for (int kk = 0; kk < 4; ++kk)
                    {
                        af[v][kk] = -Ls[ro[v] + g][c0 + 4 * kk + tg];
                        bf[v][kk] = Ls[co[v] + g][c0 + 4 * kk + tg];
                    }
                    u[v][0] = Ls[ro[v] + g][co[v] + 2 * tg];
                    u[v][1] = Ls[ro[v] + g][co[v] + 2 * tg + 1];
                }
#pragma unroll
                for (int kk = 0; kk < 4; ++kk)
#pragma unroll
                    for (int v = 0; v < 3; ++v)
                        dmma_8x8x4(u[v][0], u[v][1], af[v][kk], bf[v][kk]);
                __syncwarp();
#pragma unroll
                for (int v = 0; v < 3; ++v)
                    if (warp + 8 * v < nblk)
                    {
                        Ls[ro[v] + g][co[v] + 2 * tg] = u[v][0];
                        Ls[ro[v] + g][co[v] + 2 * tg + 1] = u[v][1];
                    }
            }
        }
    }
#else
        {   // ---- trailing update on the tensor pipe: 8x8 blocks of the lower triangle, round-robin
            const int r0 = c0 + 16, nb = nrows >> 3, nblk = nb * (nb + 1) / 2;
            for (int q = warp; q < nblk; q += 8)
            {
                int bi = (int)((sqrtf(8.0f * (float)q + 1.0f) - 1.0f) * 0.5f);
                while (bi * (bi + 1) / 2 > q) --bi;
                while ((bi + 1) * (bi + 2) / 2 <= q) ++bi;
                const int bj = q - bi * (bi + 1) / 2;
                double af[4], bf[4];
#pragma unroll
                for (int kk = 0; kk < 4; ++kk)
                {
                    af[kk] = -Ls[r0 + 8 * bi + g][c0 + 4 * kk + tg];
                    bf[kk] = Ls[r0 + 8 * bj + g][c0 + 4 * kk + tg];
                }
                double *Cp = &Ls[r0 + 8 * bi + g][r0 + 8 * bj + 2 * tg];
                double u0 = Cp[0], u1 = Cp[1];
#pragma unroll
                for (int kk = 0; kk < 4; ++kk)
                    dmma_8x8x4(u0, u1, af[kk], bf[kk]);
                Cp[0] = u0;
                Cp[1] = u1;
            }
        }
    }
#endif
    __syncthreads();
    if (d1dst)
        for (int b = 0; b < 4; ++b) d1_emit_panel(d1dst, d1tag, b, Ls, Li, tid, NT_TILE);
    TT(1);
    return *sflag;
}
#endif  // SB200_V_LOOKAHEAD

// The four 16x16 diagonal blocks of Li = L^-1 (lane j of warp b computes column j of block b); the
// strictly-upper 16x16 blocks of Li are cleared.  Ends with a block barrier.
__device__ void tile64_inv16(unsigned char *smem, int tid)
{
    double(*Ls)[LP] = reinterpret_cast<double(*)[LP]>(smem + SM_LS);
    double(*Li)[LP] = reinterpret_cast<double(*)[LP]>(smem + SM_LI);
    double *invd = reinterpret_cast<double *>(smem + SM_INVD);
    const int lane = tid & 31, warp = tid >> 5;
    if (warp < 4 && lane < 16)
    {
        const int b0 = 16 * warp, j = lane;
        double z[16];
#pragma unroll
        for (int r = 0; r < 16; ++r)
            z[r] = (r == j) ? 1.0 : 0.0;
#pragma unroll
        for (int k = 0; k < 16; ++k)
        {
            const double zk = z[k] * invd[b0 + k];
            z[k] = zk;
#pragma unroll
            for (int r = k + 1; r < 16; ++r)
                z[r] -= Ls[b0 + r][b0 + k] * zk;
        }
#pragma unroll
        for (int r = 0; r < 16; ++r)
            Li[b0 + r][b0 + j] = z[r];
    }
    else
    {   // the other threads clear the strictly-upper 16x16 blocks of Li
        const int t = tid - (tid < 128 ? 16 * (tid >> 5) + 16 : 64);    // 0..191 over the remaining threads
        for (int idx = t; idx < 6 * 256; idx += 192)
        {
            const int blk = idx >> 8, e = idx & 255;
            // upper blocks (bi < bj): (0,1) (0,2) (0,3) (1,2) (1,3) (2,3)
            const int bi = blk < 3 ? 0 : (blk < 5 ? 1 : 2);
            const int bj = blk < 3 ? blk + 1 : (blk < 5 ? blk - 1 : 3);
            Li[16 * bi + (e >> 4)][16 * bj + (e & 15)] = 0.0;
        }
    }
    __syncthreads();
}

// Off-diagonal blocks of Li from the 16x16 diagonal inverses: two levels of
// inv([A 0; C D]) = [A^-1 0; -D^-1 C A^-1  D^-1].  Ends with a block barrier.
__device__ void tile64_inv_assemble(unsigned char *smem, int tid)
{
    double(*Ls)[LP] = reinterpret_cast<double(*)[LP]>(smem + SM_LS);
    double(*Li)[LP] = reinterpret_cast<double(*)[LP]>(smem + SM_LI);
    double *Tb = reinterpret_cast<double *>(smem + SM_T);
    const int lane = tid & 31, warp = tid >> 5;
    // level 16 -> 32: two independent pairs p; T = L21 W11, X21 = -W22 T.  8 blocks of 8x8 -> 8 warps.
    {
        const int p = warp >> 2, bi = (warp >> 1) & 1, bj = warp & 1, o = 32 * p;
        dmma_block_nn<false>(&Ls[o + 16 + 8 * bi][o], LP, &Li[o][o + 8 * bj], LP, Tb + 256 * p + 8 * bi * 16 + 8 * bj, 16,
                             16, lane);
        __syncthreads();
        dmma_block_nn<true>(&Li[o + 16 + 8 * bi][o + 16], LP, Tb + 256 * p + 8 * bj, 16, &Li[o + 16 + 8 * bi][o + 8 * bj],
                            LP, 16, lane);
    }
    __syncthreads();
    // level 32 -> 64: 16 blocks of 8x8, two per warp
    for (int q = warp; q < 16; q += 8)
    {
        const int bi = q >> 2, bj = q & 3;
        dmma_block_nn<false>(&Ls[32 + 8 * bi][0], LP, &Li[0][8 * bj], LP, Tb + 8 * bi * 32 + 8 * bj, 32, 32, lane);
    }
    __syncthreads();
    for (int q = warp; q < 16; q += 8)
    {
        const int bi = q >> 2, bj = q & 3;
        dmma_block_nn<true>(&Li[32 + 8 * bi][32], LP, Tb + 8 * bj, 32, &Li[32 + 8 * bi][8 * bj], LP, 32, lane);
    }
    __syncthreads();
    TT(2);
}

// factor + full inverse (the panel-launch path and the stand-alone tile kernel)
__device__ int potrf_inv_tile64(unsigned char *smem, int tid)
{
    const int fail = potrf_tile64_factor(smem, tid);     // also leaves the 16x16 diagonal inverses in Li
    tile64_inv_assemble(smem, tid);
    return fail;
}

} // namespace sb200
