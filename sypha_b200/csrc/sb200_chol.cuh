// sb200_chol.cuh - scratch state of the dense Cholesky / triangular-solve kernels.
#pragma once
#include "sb200_common.cuh"

#include <map>
#include <utility>

namespace sb200 {

struct CholWork
{
    double *linv = nullptr;     // [T][64][64] inverses of the 64x64 diagonal blocks of L
    double *linv128 = nullptr;  // [ceil(T/2)][128][128] inverses of the 128x128 diagonal blocks
    double *gbufT = nullptr;    // the same blocks transposed (backward sweep)
    double *gbuf = nullptr;     // [T2(T2-1)/2][128][128] G_ik = W_i L_ik: the blocks the solves stream
    double2 *d1tag = nullptr;   // [T][2560] tagged hand-off of a factored diagonal tile to the next chain task
    double2 *tagged = nullptr;  // [2][ceil(T/2)*128] {value, epoch tag}: forward / backward solve hand-off
    // explicit inverse Z = L^-1 (tile count <= SB200_Z_MAX_T): the solves become two triangular GEMVs over the
    // whole GPU, x = Z'(Z b), instead of a chain of block hops (sb200_chol.cu, TASK_Z)
    double *zbuf = nullptr;     // [ld][ld] row-major, lower triangle (64x64 tiles written by the factorisation)
    double *zTbuf = nullptr;    // [ld][ld] row-major, Z' (upper triangle)
    double *ytmp = nullptr;     // [t_cap*64]
    int *ctl = nullptr;         // epochs, task counters, error flag, then the publish flags
    int2 *tasks = nullptr;      // task list of the data-flow factorisation for tasks_T tiles
    int ntasks = 0, tasks_T = 0;
    std::map<int, std::pair<int2 *, int>> task_cache;   // task lists by tile count (B&B nodes change m)
    int t_cap = 0;
    int sms = 148, potrf_occ = 1;
    int max_coop_grid = 148;
    int grid_limit = 0;         // > 0: CTAs the data-flow kernels may launch (B&B: K LPs share the GPU, and a CTA
                                // that waits for a dependency holds its SM slot; see sb200_set_concurrency_hint)
};

int chol_work_ensure(ErrorSink &err, CholWork &W, int n_pad, int n_pad_reserve = 0);
void chol_work_free(CholWork &W);
void launch_potrf(CholWork &W, int n, double *a, int ld, int *info, cudaStream_t st);
void launch_potrs(CholWork &W, int n, const double *l, int ld, double *b, cudaStream_t st);

} // namespace sb200
