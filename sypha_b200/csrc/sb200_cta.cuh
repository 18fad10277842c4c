// sb200_cta.cuh - one-CTA-per-LP Mehrotra solver (sb200_cta.cu): the throughput form of the hot path for many
// small LPs in flight (B&B node LPs, batched relaxations).
#pragma once
#include "sb200_common.cuh"
#include "sb200_kernels.cuh"

namespace sb200 {

// everything one LP needs, as device pointers (one struct per workspace, refreshed before every launch)
struct CtaLp
{
    IpmVecs V;                      // dims + iterates (V.dy == V.rhs)
    const DevParams *P;
    // model
    const int *csr_offs, *csr_inds;
    const double *csr_vals;
    const int *csc_colptr, *csc_rows;
    const double *csc_vals;
    // symbolic structure of M = A D A' (compact unit-product form)
    long long n_pairs;              // base_m (base_m + 1) / 2
    const unsigned int *chunk_ptr;
    const uint4 *term8;
    int nd;                         // doubles of d the pattern may address (pad id + 1; d[pad id] = 0)
    const double *ones;             // d of the starting point (all ones, pad slot 0)
    // B&B node rows (k_assemble_extra_rows)
    int base_m, base_n, node_k;
    const int *d_var;
    const double *d_coef;
    const int *base_colptr, *base_rows;
    const double *base_cvals;
    // pattern-only lists of the BASE model for the products (CompactLists), or nullptr: the 12-byte CSR / CSC is read
    const unsigned int *row_ptr, *col_ptr;
    const uint4 *row16, *col16;
    const double *col_sign;
    // factorisation
    double *M;                      // mpad x mpad row-major, identity pad; L in place (lower)
    int ld;
    double *linv;                   // [T][64][64] inverses of the diagonal tiles
    // warm start (sb200_node_delta.warm_start): the parent's x[warm_n] | y[warm_m] | s[warm_n], or nullptr
    const double *warm;
    int warm_n, warm_m;
    double warm_floor;
    double *export_xys;             // device: receives the final x | y | s, or nullptr
    Scalars *sc_pinned;             // pinned host mirror of the scalar block; its `done` is set last
};

static constexpr int CTA_MAX_MPAD = 2048;      // vectors of the solves live in shared memory
int cta_lp_smem_bytes();
// whether the staging of the pattern-only products (vectors + chunk pointers) fits the block's shared memory
bool cta_lists_fit(int base_m, int base_n, int node_k);
// one LP per CTA: grid = number of LPs in `lps` (device array)
int launch_ipm_cta(const CtaLp *lps, int count, cudaStream_t st);

} // namespace sb200
