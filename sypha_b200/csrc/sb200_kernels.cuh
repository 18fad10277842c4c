// sb200_kernels.cuh - launcher declarations shared by the translation units of libsypha_b200.
#pragma once
#include "sb200_common.cuh"

namespace sb200 {

// Every vector the loop touches (device addresses) - the persistent workspace of
// /root/reference/src/sypha_solver.h:76-105 (IpmWorkspace), re-laid out for the NE form.
struct IpmVecs
{
    int m, n, n_orig;
    int mpad;                 // m rounded up to SB200_TILE
    const double *c, *b;
    double *x, *y, *s;
    double *dx, *dy, *ds;
    double *resC, *resB, *resXS;
    double *d;                // x ./ s
    double *t;                // (x.*resC - resXS) ./ s
    double *rhs;              // length mpad (normal-equation right-hand side / solution in place)
    double *partial;          // [4][SB200_MAX_PARTIAL_BLOCKS]
    Scalars *sc;
    double *trace;            // [SB200_TRACE_ROWS][SB200_TRACE_COLS]
};

// Pattern-only, vector-blocked copy of a +/-1 matrix (sb200_blocked.cu): entries ordered by
// (block of the dense vector, major), 2 bytes each = 15-bit index local to the block + sign bit.
struct BlockedPattern
{
    int majors = 0, minors = 0;     // rows x cols for the A v copy, cols x rows for the A' v copy
    int nb = 0, nblk = 0;           // vector block size (doubles) and number of blocks
    long long nnz = 0;
    unsigned n_chunks = 0;          // 16-byte chunks (8 entries) stored; every segment is a whole number of chunks
    unsigned *ptr = nullptr;        // [nblk * majors + 1] segment bounds in chunks
    unsigned short *ent = nullptr;  // [8 * n_chunks]; the pad of a segment's last chunk points at the zero slot
    double *partial = nullptr;      // [nblk][majors] per-block sums (A v copy only)
};

struct CsrView
{
    int m;
    const int *offs;
    const int *inds;
    const double *vals;
    const BlockedPattern *blk = nullptr;   // host pointer; set => the blocked kernels are used
};
struct CscView
{
    int n;
    const int *colptr;
    const int *rows;
    const double *vals;
    int lanes;                // lanes per column used by the kernels (power of two <= 32)
    const BlockedPattern *blk = nullptr;
};

// CSC epilogue modes
enum
{
    CSC_PLAIN = 0,      // out = alpha*w + beta*z
    CSC_RECOVER = 1,    // ds = resC - w ; dx = (resXS - x ds)/s ; ratio test
    CSC_START_X = 2,    // x = w ; min
    CSC_START_S = 3,    // s = c - w ; min
    CSC_RESC = 4        // resC = c - s - w
};

// ---- sb200_vector.cu -------------------------------------------------------------------------
void launch_elem_min_mult(const double *x, const double *s, double *out, int n, cudaStream_t st);
void launch_corrector_rhs(const double *dx, const double *ds, double sigma, double mu, double *out,
                          int n, cudaStream_t st);
void launch_alpha_max(const double *x, const double *dx, const double *s, const double *ds, int n,
                      unsigned long long *d_ord2, double *d_result, cudaStream_t st);
void launch_reset_scalars(Scalars *sc, cudaStream_t st);
void launch_init_mu(const IpmVecs &V, const DevParams *P, cudaStream_t st);   // P: device pointer
void launch_prologue(const IpmVecs &V, cudaStream_t st);
void launch_affine_mu(const IpmVecs &V, cudaStream_t st);
void launch_corrector(const IpmVecs &V, cudaStream_t st);
void launch_affine_corrector(const IpmVecs &V, cudaStream_t st);     // both; one single-CTA kernel when n is small
void launch_update(const IpmVecs &V, const DevParams *P, cudaStream_t st);    // P: device pointer
void launch_start_shift1(const IpmVecs &V, cudaStream_t st);
void launch_start_shift2(const IpmVecs &V, cudaStream_t st);
void launch_fill(double *p, double v, long long n, cudaStream_t st);

// ---- sb200_spmv.cu ---------------------------------------------------------------------------
// out[i] = alpha * (A x)_i + beta * z[i]   (out may alias z)
void launch_spmv_csr(const CsrView &A, const double *x, const double *z, double *out, double alpha,
                     double beta, cudaStream_t st);
void launch_jacobi_diag(const CsrView &A, const double *d, double *diag, cudaStream_t st);
void launch_spmv_csc(const CscView &A, int mode, const double *v, const double *z, double *out,
                     double alpha, double beta, const IpmVecs *V, cudaStream_t st);
int pick_csc_lanes(long long nnz, int n);

// ---- sb200_blocked.cu ------------------------------------------------------------------------
int build_blocked(ErrorSink &err, int majors, int minors, long long nnz, const int *mptr, const int *midx,
                  const double *vals, int nb, bool with_partials, BlockedPattern *out, cudaStream_t st);
void free_blocked(BlockedPattern *p, cudaStream_t st = 0);
int blocked_nb_for_rows(int n);
int blocked_nb_for_cols(int m);
void launch_blk_spmv_rows(const BlockedPattern &B, const double *x, const double *z, double *out, double alpha,
                          double beta, cudaStream_t st);
void launch_blk_jacobi_diag(const BlockedPattern &B, const double *d, double *diag, cudaStream_t st);
void launch_blk_spmv_cols(const BlockedPattern &B, int mode, const double *v, const double *z, double *out,
                          double alpha, double beta, const IpmVecs *V, cudaStream_t st);

// ---- sb200_chol.cu ---------------------------------------------------------------------------
// (launch_potrf / launch_potrs: sb200_chol.cuh)
void launch_pad_identity(int n, double *a, int ld, cudaStream_t st);

// ---- sb200_assemble.cu -----------------------------------------------------------------------
struct NormalPattern
{
    int m = 0;
    long long n_pairs = 0;          // m(m+1)/2
    long long n_terms = 0;
    unsigned int *pair_ptr = nullptr;   // [n_pairs+1]
    unsigned int *term_col = nullptr;   // [n_terms] column j of each term
    double *term_w = nullptr;           // [n_terms] a_ij*a_kj, or nullptr when all products are +1
    // compact form (all products +1, n < 65535): 2-byte column ids in whole 16-byte chunks per entry,
    // padded with the id n (the workspace keeps d[n] = 0); term_col is dropped once this exists
    unsigned int *chunk_ptr = nullptr;  // [n_pairs+1], in chunks
    unsigned short *term16 = nullptr;   // [8 * n_chunks]
    unsigned int n_chunks = 0;
    int pad_id = -1;                    // the id the pad slots carry (d[pad_id] = 0)
};
// pattern-only copies of the model for the one-block solver's products (every column all +1 or all -1, ids < 65535):
// row i = the column ids of its entries, column j = the row ids, 2 bytes each in whole 16-byte chunks (pad ids: n for
// the rows, m for the columns - the product kernels keep a zero there), and the sign of every column
struct CompactLists
{
    int m = 0, n = 0;
    unsigned int *row_ptr = nullptr;    // [m+1], in chunks
    unsigned int *col_ptr = nullptr;    // [n+1], in chunks
    unsigned short *row16 = nullptr, *col16 = nullptr;
    double *col_sign = nullptr;         // [n]
    unsigned int row_chunks = 0, col_chunks = 0;
};
int build_compact_lists(ErrorSink &err, int m, int n, const int *csr_offs, const int *csr_inds, const int *csc_colptr,
                        const int *csc_rows, const double *csc_vals, CompactLists *out, cudaStream_t st);
void free_compact_lists(CompactLists *p, cudaStream_t st = 0);
int build_normal_pattern(ErrorSink &err, int m, int n, long long nnz, const int *csc_colptr,
                         const int *csc_rows, const double *csc_vals, NormalPattern *out,
                         cudaStream_t st, int pad_id = -1);
void free_normal_pattern(NormalPattern *p, cudaStream_t st = 0);
void launch_assemble_normal(const NormalPattern &P, const double *d, double *M, int ld,
                            cudaStream_t st);
void launch_syrk_dmma(int m, int k, const double *a, int lda, const double *d, double *c, int ld,
                      cudaStream_t st);
int build_csc(ErrorSink &err, int m, int n, long long nnz, const int *csr_offs, const int *csr_inds,
              const double *csr_vals, int *csc_colptr, int *csc_rows, double *csc_vals,
              cudaStream_t st);

} // namespace sb200
