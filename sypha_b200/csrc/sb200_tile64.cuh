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

static constexpr int LP = TB + 1;    // padded row stride of the tile in shared memory (doubles)

// dynamic shared-memory layout (bytes) of the kernels that factor a diagonal tile
static constexpr int SM_LS = 0;                                   // L tile (aliases the MMA staging buffers)
static constexpr int SM_MMA_BYTES = 2 * TB * KP * 8;              // 36864
static constexpr int SM_LI = SM_MMA_BYTES;                        // inverse tile
static constexpr int SM_T = SM_LI + TB * LP * 8;                  // 32x32 scratch of the inverse
static constexpr int SM_INVD = SM_T + 1024 * 8;                   // 64 reciprocal pivots
static constexpr int SM_FLAG = SM_INVD + 64 * 8;
static constexpr int SM_TOTAL = SM_FLAG + 16;
static constexpr int NT_TILE = 256;                               // threads of the tile factorisation

#ifdef SB200_TILE_TIMING
__device__ long long g_tile_timing[64];
#define TT(i) do { if (tid == 0) g_tile_timing[i] = clock64(); } while (0)
#else
#define TT(i) do { } while (0)
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

// Returns 0 or the 1-based local index of the first non-positive pivot (same value in all threads).
// On exit Ls = L (upper zeroed), Li = L^-1 (upper zero).
__device__ int potrf_inv_tile64(unsigned char *smem, int tid)
{
    TT(0);
    double(*Ls)[LP] = reinterpret_cast<double(*)[LP]>(smem + SM_LS);
    double(*Li)[LP] = reinterpret_cast<double(*)[LP]>(smem + SM_LI);
    double *Tb = reinterpret_cast<double *>(smem + SM_T);        // 32x32 scratch; also the 16x16 column buffer
    double *invd = reinterpret_cast<double *>(smem + SM_INVD);
    int *sflag = reinterpret_cast<int *>(smem + SM_FLAG);
    const int lane = tid & 31, warp = tid >> 5;
    if (tid == 0) *sflag = 0;

    for (int kb = 0; kb < 4; ++kb)
    {
        const int c0 = 16 * kb;
        __syncthreads();
        if (warp == 0)
        {   // ---- 16x16 diagonal block: lane r holds row c0+r (lanes 16..31 duplicate 0..15) ----
            // Column c of L is broadcast through shared memory (Cb); the pivot chain
            // shfl(dg) -> rsqrt -> l -> dg is issued one column ahead of the bulk update.
            const int r = lane & 15;
            double(*Cb)[17] = reinterpret_cast<double(*)[17]>(Tb);
            double a[16];
#pragma unroll
            for (int j = 0; j < 16; ++j)
                a[j] = Ls[c0 + r][c0 + j];
            double dg = a[0];
#pragma unroll
            for (int j = 1; j < 16; ++j)
                dg = (r == j) ? a[j] : dg;             // this lane's own diagonal entry
            int bad = 0;
            double d = __shfl_sync(0xffffffffu, dg, 0);
            double inv = rsqrt(d);
#pragma unroll
            for (int c = 0; c < 16; ++c)
            {
                if (!(d > 0.0) && bad == 0) bad = c0 + c + 1;
                const double l = (r == c) ? d * inv : a[c] * inv;     // L[r][c], meaningful for r >= c
                a[c] = l;
                dg -= l * l;
                Cb[c][r] = l;
                if (lane == c) invd[c0 + c] = inv;
                if (c < 15)
                {   // next pivot: long-latency chain started before this column's bulk update
                    d = __shfl_sync(0xffffffffu, dg, c + 1);
                    inv = rsqrt(d);
                }
                __syncwarp();
#pragma unroll
                for (int c2 = c + 1; c2 < 16; ++c2)
                    a[c2] -= l * Cb[c][c2];            // L[c2][c] (broadcast read)
            }
            if (lane < 16)
            {
#pragma unroll
                for (int j = 0; j < 16; ++j)
                    Ls[c0 + r][c0 + j] = (j <= r) ? a[j] : 0.0;
            }
            if (lane == 0 && bad) atomicCAS(sflag, 0, bad);
        }
        __syncthreads();
        const int nrows = 48 - c0;
        if (tid < nrows)
        {   // ---- panel: rows below, X L_dd' = A  (one thread per row) -------------------------
            const int row = c0 + 16 + tid;
            double x[16];
#pragma unroll
            for (int j = 0; j < 16; ++j)
                x[j] = Ls[row][c0 + j];
#pragma unroll
            for (int c = 0; c < 16; ++c)
            {
                const double xc = x[c] * invd[c0 + c];
                x[c] = xc;
#pragma unroll
                for (int c2 = c + 1; c2 < 16; ++c2)
                    x[c2] -= xc * Ls[c0 + c2][c0 + c];
            }
#pragma unroll
            for (int j = 0; j < 16; ++j)
                Ls[row][c0 + j] = x[j];
        }
        else if (tid >= 64 && tid < 64 + 16 && kb < 3)
        {   // meanwhile: zero the part of the block row right of the diagonal block
            const int rr = c0 + (tid - 64);
            for (int j = c0 + 16; j < TB; ++j)
                Ls[rr][j] = 0.0;
        }
        __syncthreads();
        {   // ---- trailing update on the tensor pipe: 8x8 blocks of the lower triangle, round-robin
            const int r0 = c0 + 16, nb = nrows >> 3, nblk = nb * (nb + 1) / 2;
            for (int q = warp; q < nblk; q += 8)
            {
                int bi = (int)((sqrtf(8.0f * (float)q + 1.0f) - 1.0f) * 0.5f);
                while (bi * (bi + 1) / 2 > q) --bi;
                while ((bi + 1) * (bi + 2) / 2 <= q) ++bi;
                const int bj = q - bi * (bi + 1) / 2;
                dmma_block_nt_sub(&Ls[r0 + 8 * bi][c0], &Ls[r0 + 8 * bj][c0], LP, &Ls[r0 + 8 * bi][r0 + 8 * bj], LP,
                                  16, lane);
            }
        }
    }
    __syncthreads();
    TT(1);

    // ---- inverse: four 16x16 triangular inverses, lane j of warp b computes column j -------------
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
    return *sflag;
}

} // namespace sb200
