// Per-task timeline of the data-flow Cholesky / triangular-solve kernels (globaltimer stamps).
//   nvcc -std=c++17 -O3 -DSB200_DF_TIMING -gencode arch=compute_100a,code=sm_100a \
//        -o scripts/bin/df_timeline scripts/df_timeline.cu sypha_b200/csrc/sb200_vector.cu
#ifndef SB200_DF_TIMING
#define SB200_DF_TIMING
#endif
#if !defined(SB200_TILE_TIMING) && !defined(SB200_NO_TILE_TIMING)
#define SB200_TILE_TIMING
#endif
#include "../sypha_b200/csrc/sb200_chol.cu"
#include <random>
#include <vector>
#include <algorithm>
int main(int argc, char **argv)
{
    using namespace sb200;
    const int n = argc > 1 ? atoi(argv[1]) : 1024;
    const int ld = (n + 63) / 64 * 64, T = ld / 64, T2 = (T + 1) / 2;
    std::vector<double> M((size_t)ld * ld, 0.0);
    std::mt19937 g(1);
    std::uniform_real_distribution<double> ud(-1, 1);
    for (int i = 0; i < ld; i++)
        for (int j = 0; j <= i; j++)
            M[(size_t)i * ld + j] = (i == j) ? ld * 1.0 + 1.0 : ud(g);
    double *dA, *dB;
    int *info;
    cudaMalloc(&dA, sizeof(double) * ld * ld);
    cudaMalloc(&dB, sizeof(double) * ld);
    cudaMalloc(&info, 4);
    cudaMemset(info, 0, 4);
    std::vector<double> b(ld, 1.0);
    ErrorSink err;
    CholWork W;
    if (chol_work_ensure(err, W, ld)) { printf("ensure failed: %s\n", err.msg.c_str()); return 1; }
    cudaEvent_t e0, e1, e2;
    cudaEventCreate(&e0); cudaEventCreate(&e1); cudaEventCreate(&e2);
    for (int rep = 0; rep < 4; rep++)
    {
        cudaMemcpy(dA, M.data(), sizeof(double) * ld * ld, cudaMemcpyHostToDevice);
        cudaMemcpy(dB, b.data(), sizeof(double) * ld, cudaMemcpyHostToDevice);
        cudaDeviceSynchronize();
        cudaEventRecord(e0);
        launch_potrf(W, n, dA, ld, info, 0);
        cudaEventRecord(e1);
        launch_potrs(W, n, dA, ld, dB, 0);
        cudaEventRecord(e2);
        cudaEventSynchronize(e2);
        float ms1, ms2;
        cudaEventElapsedTime(&ms1, e0, e1);
        cudaEventElapsedTime(&ms2, e1, e2);
        printf("rep %d: potrf %.1f us, potrs %.1f us (%s)\n", rep, ms1 * 1000, ms2 * 1000, cudaGetErrorString(cudaGetLastError()));
    }
    static unsigned long long tm[8192][12];
    cudaMemcpyFromSymbol(tm, g_df_time, sizeof tm);
    std::vector<int2> tasks(W.ntasks);
    cudaMemcpy(tasks.data(), W.tasks, sizeof(int2) * W.ntasks, cudaMemcpyDeviceToHost);
    unsigned long long t0 = ~0ull;
    for (int t = 0; t < W.ntasks && t < 4096; t++) t0 = std::min(t0, tm[t][0]);
    printf("# potrf tasks: type i j claim acc_done d1 done (us from first claim)\n");
    for (int t = 0; t < W.ntasks && t < 4096; t++)
    {
        const int type = tasks[t].x >> 16, i = tasks[t].x & 0xffff, j = tasks[t].y;
        const bool diag = type == 2;
        if (T > 20 && !(diag || type == 1 || i == j + 2)) continue;
        printf("%s %3d %3d  %8.2f %8.2f %8.2f %8.2f\n", type == 1 ? "pair" : (diag ? "chain" : (type == 4 ? "zinv" : "tile")), i, j,
               (tm[t][0] - t0) * 1e-3, type == 1 ? 0.0 : (tm[t][1] - t0) * 1e-3, diag ? (tm[t][2] - t0) * 1e-3 : 0.0,
               (tm[t][3] - t0) * 1e-3);
        if (diag && j > 0)
            printf("        D1(j-1) seen %8.2f  Ljj loaded %8.2f  subst done %8.2f  | stores issued %8.2f\n",
                   (tm[t][8] - t0) * 1e-3, (tm[t][9] - t0) * 1e-3, (tm[t][10] - t0) * 1e-3, (tm[t][11] - t0) * 1e-3);
        if (diag && j > 0)
            printf("        trsm+publish done %8.2f  update done %8.2f  factor done %8.2f  inv16 done %8.2f\n",
                   (tm[t][4] - t0) * 1e-3, (tm[t][5] - t0) * 1e-3, (tm[t][6] - t0) * 1e-3, (tm[t][7] - t0) * 1e-3);
    }
    unsigned long long s0 = ~0ull;
    for (int t = 0; t < 2 * T2; t++) s0 = std::min(s0, tm[4096 + t][0]);
    printf("# trsv tasks: t claim acc_done done (us)\n");
    for (int t = 0; t < 2 * T2 && t < 64; t++)
    {
        printf("trsv %3d  %8.2f %8.2f %8.2f", t, (tm[4096 + t][0] - s0) * 1e-3, (tm[4096 + t][1] - s0) * 1e-3,
               (tm[4096 + t][3] - s0) * 1e-3);
        if (t >= T2) printf("   last-x seen %8.2f  y_i seen %8.2f  W done %8.2f", (tm[4096 + t][4] - s0) * 1e-3,
                            (tm[4096 + t][5] - s0) * 1e-3, (tm[4096 + t][6] - s0) * 1e-3);
        printf("\n");
    }
#ifdef SB200_TILE_TIMING
    {
        long long tt[64];
        cudaMemcpyFromSymbol(tt, g_tile_timing, sizeof tt);
        printf("# last tile factorisation (cycles): total %lld\n", tt[1] - tt[0]);
#if SB200_V_LOOKAHEAD
        for (int kb = 0; kb < 4; kb++)
            printf("   step %d (cycles from step start): warps 1-7 done %lld  emitted %lld  P done %lld  N done %lld  step end %lld\n", kb,
                   tt[9 + 4 * kb] - tt[8 + 4 * kb], tt[24 + kb] - tt[8 + 4 * kb], tt[10 + 4 * kb] - tt[8 + 4 * kb],
                   kb < 3 ? tt[11 + 4 * kb] - tt[8 + 4 * kb] : 0ll, (kb < 3 ? tt[12 + 4 * kb] : tt[1]) - tt[8 + 4 * kb]);
#else
        for (int kb = 0; kb < 4; kb++)
            printf("   panel %d: diag16 %lld  rows-below %lld  trailing %lld\n", kb, tt[9 + 4 * kb] - tt[8 + 4 * kb],
                   tt[10 + 4 * kb] - tt[9 + 4 * kb], (kb < 3 ? tt[12 + 4 * kb] : tt[1]) - tt[10 + 4 * kb]);
#endif
    }
#endif
    int hinfo = -1;
    cudaMemcpy(&hinfo, info, 4, cudaMemcpyDeviceToHost);
    std::vector<double> x(ld);
    cudaMemcpy(x.data(), dB, sizeof(double) * ld, cudaMemcpyDeviceToHost);
    // residual of M x = b
    double rmax = 0;
    for (int i = 0; i < n; i++)
    {
        double s = 0;
        for (int j = 0; j < n; j++) s += (j <= i ? M[(size_t)i * ld + j] : M[(size_t)j * ld + i]) * x[j];
        rmax = fmax(rmax, fabs(s - 1.0));
    }
    printf("info %d residual %.3e\n", hinfo, rmax);
    return 0;
}
