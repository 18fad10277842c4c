// scripts/dmma_occupancy.cu - DMMA (mma.sync.m8n8k4.f64) rate of one SM against the number of resident warps and the
// number of independent accumulator chains per warp: what a one-CTA-per-SM kernel (sb200_cta.cu) can expect.
#include <cstdio>
#include <cuda_runtime.h>
template <int ILP> __global__ void k(double *out, double x, double y, int trips)
{
    double c[ILP][2];
    for (int j = 0; j < ILP; j++) { c[j][0] = x + j; c[j][1] = y; }
    const double a = x + (threadIdx.x & 3), b = y + (threadIdx.x >> 2);
    for (int t = 0; t < trips; t++)
#pragma unroll
        for (int i = 0; i < 8; i++)
#pragma unroll
            for (int j = 0; j < ILP; j++)
                asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                             : "+d"(c[j][0]), "+d"(c[j][1]) : "d"(a), "d"(b));
    double s = 0;
    for (int j = 0; j < ILP; j++) s += c[j][0] + c[j][1];
    out[threadIdx.x + (size_t)blockIdx.x * blockDim.x] = s;
}
template <int ILP> static void run(int ctas_per_sm, int threads, double *out)
{
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int trips = 4096;
    k<ILP><<<148 * ctas_per_sm, threads>>>(out, 1.0, 0.5, trips);
    cudaEventRecord(e0); k<ILP><<<148 * ctas_per_sm, threads>>>(out, 1.0, 0.5, trips); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double flops = 148.0 * ctas_per_sm * (threads / 32) * trips * 8.0 * ILP * 512.0;
    printf("warps/SM %2d, chains/warp %d: %.2f TFLOP/s (%.1f %% of 37.2)\n", ctas_per_sm * threads / 32, ILP, flops / ms / 1e9, flops / ms / 1e9 / 37.2 * 100);
}
int main()
{
    double *out; cudaMalloc(&out, 8 * 148 * 2 * 1024);
    run<8>(1, 128, out); run<8>(1, 256, out); run<8>(1, 512, out); run<8>(1, 1024, out); run<8>(2, 512, out);
    run<4>(1, 512, out); run<2>(1, 512, out); run<1>(1, 512, out); run<16>(1, 512, out);
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
}
