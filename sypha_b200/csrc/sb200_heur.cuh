// sb200_heur.cuh - launcher of the per-node branching / incumbent kernel (sb200_heur.cu).
#pragma once
#include "sb200_common.cuh"

namespace sb200 {

struct HeurArgs
{
    int m0, n0;                       // base rows, original (non-slack) columns
    const int *row_ptr, *row_cols;    // CSR of the node model (rows >= m0 and columns >= n0 are skipped)
    const int *col_ptr, *col_rows;    // CSC of the node model
    const double *c;
    const double *x_lp;               // the LP point, where the solver left it
    int k;                            // the node's decisions (coef < 0: variable fixed to 0)
    const int *var;
    const double *coef;
    int *list, *sorted;               // scratch, n0 ints each
    unsigned char *cover_x;           // out: the cover, n0 bytes
    sb200_heur_result *out;           // out (device)
    // the reference's rules (SB200_HEUR_REFERENCE) also read:
    int rules, branch_rule;           // SB200_HEUR_*, SB200_BRANCH_*
    double tol;                       // integrality tolerance (kBnbIntegralityTol = 1e-6)
    const double *y_lp;               // dual point (dual guidance)
    const double *b;                  // right-hand sides (base rows must be 1)
    const double *row_vals, *col_vals; // CSR / CSC values (base entries of original columns must be 1)
    unsigned char *nif_x;             // out: the NearestIntegerFixing rounding, n0 bytes
    double *score;                    // scratch, n0 doubles: cached repair scores
    volatile int *host_flag;          // pinned host word set to flag_value (behind a system fence) when `out` is complete
    int flag_value;
};

size_t heur_smem_bytes(int m0, int n0, int rules);
int launch_node_heuristics(const HeurArgs &a, cudaStream_t st);
// the reference-rules kernel for `count` nodes of the same base model in one launch (d_args: device array)
int launch_node_heuristics_batch(const HeurArgs *d_args, int count, int m0, int n0, cudaStream_t st);

} // namespace sb200
