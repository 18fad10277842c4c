// Micro-benchmark of the 64x64 tile factorisation (+inverse) that sits on the Cholesky critical path.
//   nvcc -std=c++17 -O3 -DSB200_TILE_TIMING -gencode arch=compute_100a,code=sm_100a \
//        -o scripts/bin/tile_microbench scripts/tile_microbench.cu sypha_b200/csrc/sb200_vector.cu
#ifndef SB200_TILE_TIMING
#define SB200_TILE_TIMING
#endif
#include "../sypha_b200/csrc/sb200_chol.cu"
#include <random>
#include <vector>
int main()
{
    using namespace sb200;
    const int ld = 64;
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
    double *dA, *dL;
    int *info;
    cudaMalloc(&dA, 64 * 64 * 8);
    cudaMalloc(&dL, 64 * 64 * 8);
    cudaMalloc(&info, 4);
    cudaMemset(info, 0, 4);
    cudaFuncSetAttribute(k_potrf_first, cudaFuncAttributeMaxDynamicSharedMemorySize, SM_TOTAL);
    for (int rep = 0; rep < 3; rep++)
    {
        cudaMemcpy(dA, M.data(), 64 * 64 * 8, cudaMemcpyHostToDevice);
        cudaEvent_t e0, e1;
        cudaEventCreate(&e0);
        cudaEventCreate(&e1);
        cudaEventRecord(e0);
        k_potrf_first<<<1, NT_TILE, SM_TOTAL>>>(dA, ld, dL, info);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        long long t[64];
        cudaMemcpyFromSymbol(t, g_tile_timing, sizeof t);
        printf("rep %d: kernel %.2f us; factor %lld cycles, inverse %lld cycles\n", rep, ms * 1000, t[1] - t[0], t[2] - t[1]);
        for (int kb = 0; kb < 4; kb++)
            printf("   panel %d: diag16 %lld  rows-below %lld  trailing %lld\n", kb, t[9 + 4 * kb] - t[8 + 4 * kb],
                   t[10 + 4 * kb] - t[9 + 4 * kb], (kb < 3 ? t[12 + 4 * kb] : t[1]) - t[10 + 4 * kb]);
    }
    std::vector<double> L(64 * 64), Li(64 * 64);
    cudaMemcpy(L.data(), dA, 64 * 64 * 8, cudaMemcpyDeviceToHost);
    cudaMemcpy(Li.data(), dL, 64 * 64 * 8, cudaMemcpyDeviceToHost);
    double err = 0, ierr = 0;
    for (int i = 0; i < 64; i++)
        for (int j = 0; j <= i; j++)
        {
            double s = 0;
            for (int k = 0; k <= j; k++) s += L[i * 64 + k] * L[j * 64 + k];
            err = fmax(err, fabs(s - M[i * 64 + j]));
        }
    for (int i = 0; i < 64; i++)
        for (int j = 0; j < 64; j++)
        {
            double s = 0;
            for (int k = 0; k < 64; k++) s += ((k <= i) ? L[i * 64 + k] : 0.0) * Li[k * 64 + j];
            ierr = fmax(ierr, fabs(s - (i == j)));
        }
    int hinfo = -1;
    cudaMemcpy(&hinfo, info, 4, cudaMemcpyDeviceToHost);
    printf("LL' err %.3e  L*Linv-I err %.3e info %d cuda=%s\n", err, ierr, hinfo, cudaGetErrorString(cudaGetLastError()));
    return 0;
}
