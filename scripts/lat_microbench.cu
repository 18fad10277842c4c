// Latency / issue-rate probes that shape the Cholesky critical path (one CTA, one SM).
//   nvcc -std=c++17 -O3 -gencode arch=compute_100a,code=sm_100a -o scripts/bin/lat_microbench scripts/lat_microbench.cu
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void dmma(double &c0, double &c1, double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__global__ void probe(double *out, long long *cyc, double seed)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    __shared__ double sm[1024];
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) sm[i] = seed + i * 1e-9;
    __syncthreads();
    long long t0, t1;
    double x = seed + lane * 1e-6, y = 1.0000001, z = 0.5;
    const int N = 256;
    // 0: dependent DFMA
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; ++i) x = fma(x, y, z);
    t1 = clock64();
    if (threadIdx.x == 0) cyc[0] = (t1 - t0);
    // 1: dependent DMUL
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; ++i) x = x * y;
    t1 = clock64();
    if (threadIdx.x == 0) cyc[1] = (t1 - t0);
    // 2: dependent rsqrt(double)
    x = fabs(x) + 2.0;
    t0 = clock64();
#pragma unroll 4
    for (int i = 0; i < N; ++i) x = rsqrt(x) + 1.5;
    t1 = clock64();
    if (threadIdx.x == 0) cyc[2] = (t1 - t0);
    // 3: dependent 64-bit shuffle
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; ++i) x = __shfl_sync(0xffffffffu, x, (lane + 1) & 31);
    t1 = clock64();
    if (threadIdx.x == 0) cyc[3] = (t1 - t0);
    // 4: dependent DMMA chain (accumulator dependency)
    double c0 = x, c1 = y;
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; ++i) dmma(c0, c1, y, z);
    t1 = clock64();
    if (threadIdx.x == 0) cyc[4] = (t1 - t0);
    // 5: 4 independent DMMA chains per warp
    double d0 = x, d1 = y, e0 = x, e1 = z, f0 = y, f1 = z;
    t0 = clock64();
#pragma unroll 4
    for (int i = 0; i < N; ++i)
    {
        dmma(c0, c1, y, z);
        dmma(d0, d1, y, z);
        dmma(e0, e1, y, z);
        dmma(f0, f1, y, z);
    }
    t1 = clock64();
    if (threadIdx.x == 0) cyc[5] = (t1 - t0);
    // 6: dependent LDS (pointer chase through doubles)
    int idx = lane;
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; ++i) idx = ((int)sm[idx & 1023] + idx + 1) & 1023;
    t1 = clock64();
    if (threadIdx.x == 0) cyc[6] = (t1 - t0);
    // 7: __syncthreads round trip
    t0 = clock64();
    for (int i = 0; i < 64; ++i) __syncthreads();
    t1 = clock64();
    if (threadIdx.x == 0) cyc[7] = (t1 - t0) * 4;
    // 8: approx rsqrt + 2 Newton steps
    x = fabs(x) + 2.0;
    t0 = clock64();
#pragma unroll 4
    for (int i = 0; i < N; ++i)
    {
        double r;
        asm volatile("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
        double h = 0.5 * x * r, e = fma(-h, r, 0.5);
        r = fma(r, e, r);
        h = 0.5 * x * r;
        e = fma(-h, r, 0.5);
        r = fma(r, e, r);
        x = r + 1.5;
    }
    t1 = clock64();
    if (threadIdx.x == 0) cyc[8] = (t1 - t0);
    // 9: dependent DADD
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; ++i) x = x + y;
    t1 = clock64();
    if (threadIdx.x == 0) cyc[9] = (t1 - t0);
    out[threadIdx.x] = x + c0 + c1 + d0 + d1 + e0 + e1 + f0 + f1 + idx + warp;
}
int main()
{
    double *out;
    long long *cyc, h[16];
    cudaMalloc(&out, 8 * 1024);
    cudaMalloc(&cyc, 8 * 16);
    const char *names[] = {"dep DFMA", "dep DMUL", "dep rsqrt(double)+add", "dep shfl64", "dep DMMA", "4 indep DMMA chains (per 4)",
                           "dep LDS+int ops", "__syncthreads", "rsqrt.approx+2NR+add", "dep DADD"};
    for (int threads : {32, 128, 256})
    {
        for (int rep = 0; rep < 2; ++rep)
        {
            probe<<<1, threads>>>(out, cyc, 1.25);
            cudaDeviceSynchronize();
        }
        cudaMemcpy(h, cyc, sizeof h, cudaMemcpyDeviceToHost);
        printf("threads %d (%s)\n", threads, cudaGetErrorString(cudaGetLastError()));
        for (int i = 0; i < 10; ++i) printf("  %-32s %8.1f cycles/op\n", names[i], h[i] / 256.0);
    }
    return 0;
}
