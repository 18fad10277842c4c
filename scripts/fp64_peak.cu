// scripts/fp64_peak.cu - the FP64 denominators bench.py's roofline uses, measured on the box it runs on.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a scripts/fp64_peak.cu -lcublas -o scripts/bin/fp64_peak
// Prints one JSON object: sustained DMMA issue rate (mma.sync.m8n8k4.f64, SASS DMMA.8x8x4) and DFMA rate of the whole
// chip over >= 5 ms kernels, and cuBLAS DGEMM at 4096^3 and 8192^3 (SURVEY.md 8d asks for the latter as the
// FP64 tensor denominator; MEASURED_PEAKS.json carries no FP64 entry).
#include <cstdio>
#include <vector>
#include <cuda_runtime.h>
#include <cublas_v2.h>

template <int ILP> __global__ void k_dmma(double *out, double x, double y, int trips)
{
    double c[ILP][2];
    for (int j = 0; j < ILP; j++) { c[j][0] = x + j; c[j][1] = y; }
    const double a = x + (threadIdx.x & 3), b = y + (threadIdx.x >> 2);
    for (int t = 0; t < trips; t++)
    {
#pragma unroll
        for (int i = 0; i < 16; i++)
#pragma unroll
            for (int j = 0; j < ILP; j++)
                asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                             : "+d"(c[j][0]), "+d"(c[j][1]) : "d"(a), "d"(b));
    }
    double s = 0;
    for (int j = 0; j < ILP; j++) s += c[j][0] + c[j][1];
    out[threadIdx.x + (size_t)blockIdx.x * blockDim.x] = s;
}

template <int ILP> __global__ void k_dfma(double *out, double x, double y, int trips)
{
    double a[ILP];
    for (int j = 0; j < ILP; j++) a[j] = x + threadIdx.x + j;
    for (int t = 0; t < trips; t++)
    {
#pragma unroll
        for (int i = 0; i < 16; i++)
#pragma unroll
            for (int j = 0; j < ILP; j++) a[j] = fma(a[j], y, x);
    }
    double s = 0;
    for (int j = 0; j < ILP; j++) s += a[j];
    out[threadIdx.x + (size_t)blockIdx.x * blockDim.x] = s;
}

template <class F> static float best_ms(F launch, int reps)
{
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    launch();
    cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < reps; r++)
    {
        cudaEventRecord(e0);
        launch();
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    return best;
}

int main()
{
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, 0);
    const int sms = prop.multiProcessorCount;
    double *out;
    cudaMalloc(&out, sizeof(double) * (size_t)sms * 8 * 1024);
    const int trips = 4096;
    // 16 warps per CTA, 2 CTAs per SM: 8 independent DMMA chains per warp
    const float ms_dmma = best_ms([&] { k_dmma<8><<<sms * 2, 512>>>(out, 1.0, 0.5, trips); }, 5);
    const double dmma_flops = (double)sms * 2 * 16 * trips * 16 * 8 * (2.0 * 8 * 8 * 4);
    const float ms_dfma = best_ms([&] { k_dfma<8><<<sms * 2, 1024>>>(out, 1.0, 0.999, trips); }, 5);
    const double dfma_flops = (double)sms * 2 * 1024 * trips * 16 * 8 * 2.0;

    cublasHandle_t h;
    cublasCreate(&h);
    double dg[2] = {0, 0};
    const int sizes[2] = {4096, 8192};
    for (int s = 0; s < 2; s++)
    {
        const int n = sizes[s];
        double *A, *B, *Cm;
        cudaMalloc(&A, sizeof(double) * n * n);
        cudaMalloc(&B, sizeof(double) * n * n);
        cudaMalloc(&Cm, sizeof(double) * n * n);
        std::vector<double> hst((size_t)n * n);
        for (size_t i = 0; i < hst.size(); i++) hst[i] = (double)((i * 2654435761u) % 1000) / 1000.0 - 0.5;
        cudaMemcpy(A, hst.data(), sizeof(double) * n * n, cudaMemcpyHostToDevice);
        cudaMemcpy(B, hst.data(), sizeof(double) * n * n, cudaMemcpyHostToDevice);
        const double one = 1.0, zero = 0.0;
        const float ms = best_ms([&] { cublasDgemm(h, CUBLAS_OP_N, CUBLAS_OP_T, n, n, n, &one, A, n, B, n, &zero, Cm, n); }, 5);
        dg[s] = 2.0 * n * n * n / ms / 1e9;
        cudaFree(A);
        cudaFree(B);
        cudaFree(Cm);
    }
    int clk = 0;
    cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    printf("{\"device\": \"%s\", \"sms\": %d, \"sm_clock_max_mhz\": %.0f, \"dmma_tflops\": %.2f, \"dmma_ms\": %.3f, "
           "\"dfma_tflops\": %.2f, \"dfma_ms\": %.3f, \"cublas_dgemm_4096_tflops\": %.2f, \"cublas_dgemm_8192_tflops\": %.2f, "
           "\"cuda_error\": \"%s\"}\n",
           prop.name, sms, clk / 1e3, dmma_flops / ms_dmma / 1e9, ms_dmma, dfma_flops / ms_dfma / 1e9, ms_dfma, dg[0], dg[1],
           cudaGetErrorString(cudaGetLastError()));
    return 0;
}
