/* oracle/refbuild/gsl/gsl_matrix.h - TEST INFRASTRUCTURE, not product code.
 *
 * GSL is not installed in this image and there is no network.  The reference's CUDA solver uses GSL
 * only for its host-side starting point (src/sypha_solver_init.cpp:543-652) and three scalar helpers
 * (src/sypha_solver.cpp:600-601,622,697-698): 17 symbols in all.  This header and gsl_shim.cpp provide
 * exactly those symbols with GSL's published semantics (row-major storage with a row stride `tda`,
 * which the reference pokes directly, init.cpp:570-611) so that the UNMODIFIED reference sources compile
 * and run as the side-by-side baseline (oracle/Makefile -> oracle/_ref/).  Nothing under sypha_b200/
 * includes or links this.
 */
#ifndef SB200_REFBUILD_GSL_MATRIX_H
#define SB200_REFBUILD_GSL_MATRIX_H
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct { size_t size, stride; double *data; void *block; int owner; } gsl_vector;
typedef struct { size_t size1, size2, tda; double *data; void *block; int owner; } gsl_matrix;
typedef struct { size_t size; size_t *data; } gsl_permutation;

typedef enum { CblasNoTrans = 111, CblasTrans = 112, CblasConjTrans = 113 } CBLAS_TRANSPOSE_t;

gsl_vector *gsl_vector_alloc(size_t n);
void gsl_vector_free(gsl_vector *v);
double gsl_vector_min(const gsl_vector *v);
int gsl_vector_add_constant(gsl_vector *v, double x);

gsl_matrix *gsl_matrix_calloc(size_t n1, size_t n2);
void gsl_matrix_free(gsl_matrix *m);

gsl_permutation *gsl_permutation_alloc(size_t n);
void gsl_permutation_free(gsl_permutation *p);

int gsl_blas_dgemm(CBLAS_TRANSPOSE_t ta, CBLAS_TRANSPOSE_t tb, double alpha, const gsl_matrix *A,
                   const gsl_matrix *B, double beta, gsl_matrix *C);
int gsl_blas_dgemv(CBLAS_TRANSPOSE_t ta, double alpha, const gsl_matrix *A, const gsl_vector *x, double beta,
                   gsl_vector *y);
int gsl_blas_ddot(const gsl_vector *x, const gsl_vector *y, double *result);

int gsl_linalg_LU_decomp(gsl_matrix *A, gsl_permutation *p, int *signum);
int gsl_linalg_LU_invert(const gsl_matrix *LU, const gsl_permutation *p, gsl_matrix *inverse);

#ifdef __cplusplus
}
#endif

/* gsl_math.h / gsl_minmax.h / gsl_pow_int.h */
#define GSL_MAX(a, b) ((a) > (b) ? (a) : (b))
#define GSL_MIN(a, b) ((a) < (b) ? (a) : (b))
static inline double gsl_max(double a, double b) { return GSL_MAX(a, b); }
static inline double gsl_min(double a, double b) { return GSL_MIN(a, b); }
static inline double gsl_pow_3(double x) { return x * x * x; }

#endif
