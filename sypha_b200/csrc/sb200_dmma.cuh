// sb200_dmma.cuh - FP64 tensor-core (DMMA) tile primitives shared by the Cholesky trailing update
// and the SYRK assembly kernel.
#pragma once
#include "sb200_common.cuh"

namespace sb200 {

static constexpr int TB = 64;
static constexpr int KC = 32;        // K chunk staged in shared memory
static constexpr int KP = KC + 4;    // padded row stride (doubles)

// ---------------------------------------------------------------------------------------------
// FP64 tensor-core MMA  D(8x8) = A(8x4) B(4x8) + C
//   A: lane holds A[lane/4][lane%4];  B: lane holds B[lane%4][lane/4];
//   C: lane holds C[lane/4][2*(lane%4)+{0,1}]
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void dmma_8x8x4(double &c0, double &c1, double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

// acc((MI*8) x (NJ*8) warp tile at row0/col0) += sign * As(rows row0.., KC) * Bs(rows col0.., KC)'
template <int MI, int NJ>
__device__ __forceinline__ void warp_mma(const double (*As)[KP], const double (*Bs)[KP], int row0, int col0,
                                         int lane, double sign, double acc[MI][NJ][2])
{
    const int g = lane >> 2, tg = lane & 3;
#pragma unroll
    for (int kk = 0; kk < KC; kk += 4)
    {
        double a[MI], b[NJ];
#pragma unroll
        for (int i = 0; i < MI; ++i)
            a[i] = sign * As[row0 + i * 8 + g][kk + tg];
#pragma unroll
        for (int j = 0; j < NJ; ++j)
            b[j] = Bs[col0 + j * 8 + g][kk + tg];
#pragma unroll
        for (int i = 0; i < MI; ++i)
#pragma unroll
            for (int j = 0; j < NJ; ++j)
                dmma_8x8x4(acc[i][j][0], acc[i][j][1], a[i], b[j]);
    }
}
__device__ __forceinline__ void warp_mma_32x32(const double (*As)[KP], const double (*Bs)[KP], int wm,
                                               int wn, int lane, double sign, double acc[4][4][2])
{
    warp_mma<4, 4>(As, Bs, wm * 32, wn * 32, lane, sign, acc);
}

// cooperative load of a 64 x KC block (row-major, leading dim ld) into padded shared memory,
// optionally scaling column kk by scale[kk] (SYRK: A diag(d))
__device__ __forceinline__ void load_tile_64xKC(double (*S)[KP], const double *__restrict__ g, size_t ld,
                                                int tid, int nthreads, const double *scale)
{
    for (int idx = tid; idx < TB * (KC / 2); idx += nthreads)
    {
        const int r = idx / (KC / 2), c2 = (idx % (KC / 2)) * 2;
        double2 v = *reinterpret_cast<const double2 *>(g + (size_t)r * ld + c2);
        if (scale)
        {
            v.x *= scale[c2];
            v.y *= scale[c2 + 1];
        }
        *reinterpret_cast<double2 *>(&S[r][c2]) = v;
    }
}


} // namespace sb200
