// Isolated timing of the 16x16 diagonal-block chain (one warp), repeated to separate cold
// instruction-cache effects from the dependent-latency chain.
//   nvcc -std=c++17 -O3 -gencode arch=compute_100a,code=sm_100a -o scripts/bin/diag16_bench scripts/diag16_bench.cu
#include <cstdio>
#include <cuda_runtime.h>
#include <vector>
#include <random>
__device__ __forceinline__ double fast_rsqrt(double x)
{
    double r;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    double h = 0.5 * x * r, e = fma(-h, r, 0.5);
    r = fma(r, e, r);
    h = 0.5 * x * r;
    e = fma(-h, r, 0.5);
    return fma(r, e, r);
}
template <int V>
__device__ __forceinline__ void diag16(double (*Ls)[65], double (*Li)[65], double (*Cb)[17], int c0, int lane)
{
    const int r = lane & 15;
    const bool inv_lane = lane >= 16;
    double a[16];
#pragma unroll
    for (int j = 0; j < 16; ++j)
        a[j] = inv_lane ? (j == r ? 1.0 : 0.0) : Ls[c0 + r][c0 + j];
    double dg = a[0];
#pragma unroll
    for (int j = 1; j < 16; ++j)
        dg = (r == j) ? a[j] : dg;
    double d = __shfl_sync(0xffffffffu, dg, 0);
    double inv = V == 1 ? fast_rsqrt(d) : rsqrt(d);
#pragma unroll
    for (int c = 0; c < 16; ++c)
    {
        const double l = (!inv_lane && r == c) ? d * inv : a[c] * inv;
        a[c] = l;
        dg -= l * l;
        if (!inv_lane) Cb[c][r] = l;
        if (c < 15)
        {
            d = __shfl_sync(0xffffffffu, dg, c + 1);
            inv = V == 1 ? fast_rsqrt(d) : rsqrt(d);
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
            Cb[j][16] = a[j];
    }
}
template <int V>
__global__ void bench(const double *M, long long *cyc, double *out)
{
    __shared__ double Ls[64][65], Cb[16][17];
    double(*Li)[65] = Ls;   // inverse lanes write strictly below-diagonal-free positions; aliasing is fine for timing
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int rep = 0; rep < 8; ++rep)
    {
        for (int idx = tid; idx < 4096; idx += blockDim.x) Ls[idx >> 6][idx & 63] = M[idx];
        __syncthreads();
        long long t0 = clock64();
        if (warp == 0) diag16<V>(Ls, Li, Cb, 16 * (rep & 3), lane);
        __syncthreads();
        long long t1 = clock64();
        if (tid == 0) cyc[rep] = t1 - t0;
    }
    out[tid] = Ls[tid & 63][tid >> 6] + Li[tid & 63][tid >> 6];
}
int main()
{
    std::vector<double> B(64 * 72), M(64 * 64);
    std::mt19937 g(1);
    std::normal_distribution<double> nd;
    for (auto &v : B) v = nd(g);
    for (int i = 0; i < 64; i++)
        for (int j = 0; j < 64; j++)
        {
            double s = 0;
            for (int k = 0; k < 72; k++) s += B[i * 72 + k] * B[j * 72 + k];
            M[i * 64 + j] = s + (i == j ? 0.1 : 0);
        }
    double *dM, *out;
    long long *cyc, h[8];
    cudaMalloc(&dM, 4096 * 8);
    cudaMalloc(&out, 256 * 8);
    cudaMalloc(&cyc, 64);
    cudaMemcpy(dM, M.data(), 4096 * 8, cudaMemcpyHostToDevice);
    for (int v = 0; v < 2; ++v)
        for (int threads : {32, 256})
        {
            if (v == 0) bench<0><<<1, threads>>>(dM, cyc, out);
            else bench<1><<<1, threads>>>(dM, cyc, out);
            cudaDeviceSynchronize();
            cudaMemcpy(h, cyc, 64, cudaMemcpyDeviceToHost);
            printf("variant %d threads %d: cycles per 16x16 block:", v, threads);
            for (int i = 0; i < 8; ++i) printf(" %lld", h[i]);
            printf("  (%s)\n", cudaGetErrorString(cudaGetLastError()));
        }
    return 0;
}
