// sb200_chol.cuh - scratch state of the dense Cholesky / triangular-solve kernels.
#pragma once
#include "sb200_common.cuh"

namespace sb200 {

struct CholWork
{
    double *linv = nullptr;     // [T][64][64] inverses of the diagonal blocks of L
    int *ctl = nullptr;         // epochs, exit tickets, error flag, then 2*T publish flags
    int t_cap = 0;
    int max_coop_grid = 148;
};

int chol_work_ensure(ErrorSink &err, CholWork &W, int n_pad);
void chol_work_free(CholWork &W);
void launch_potrf(CholWork &W, int n, double *a, int ld, int *info, cudaStream_t st);
void launch_potrs(CholWork &W, int n, const double *l, int ld, double *b, cudaStream_t st);

} // namespace sb200
