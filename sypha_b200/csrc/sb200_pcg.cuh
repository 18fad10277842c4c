// sb200_pcg.cuh - state of the Jacobi-PCG normal-equations solver (replaces KrylovSolveWorkspace,
// /root/reference/src/sypha_solver_krylov.h:9-36).
#pragma once
#include "sb200_kernels.cuh"

namespace sb200 {

struct PcgVecs
{
    int m;
    const double *rhs;      // [m]
    const double *diag;     // [m] Jacobi diagonal of A D A'
    double *x;              // [m] solution
    double *r, *z, *p, *Ap; // [m]
    double *q;              // [n]
    const double *dscale;   // [n] D, or nullptr for D = I (starting point)
    double *partial;        // [2][SB200_MAX_PARTIAL_BLOCKS]
};

void launch_cg_init(const PcgVecs &C, Scalars *sc, const DevParams *P, double fixed_tol, int honour_done,
                    cudaStream_t st);
void launch_cg_iteration(const CsrView &A, const CscView &At, const PcgVecs &C, const IpmVecs &V,
                         const DevParams *P, int max_iter_override, cudaStream_t st);
void launch_cg_check(Scalars *sc, cudaStream_t st);
// sb200_blocked.cu: the two products of a CG iteration on the blocked pattern
void launch_blk_cg_matvec(const BlockedPattern &B, const PcgVecs &C, Scalars *sc, cudaStream_t st);
void launch_blk_cg_cols(const BlockedPattern &B, const double *p, double *q, const double *dscale, const Scalars *sc,
                        cudaStream_t st);
void launch_spmv_csc_cg(const CscView &A, const double *p, double *q, const double *dscale,
                        const Scalars *sc, cudaStream_t st);

} // namespace sb200
