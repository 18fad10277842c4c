// oracle/refbuild/gsl_shim.cpp - TEST INFRASTRUCTURE, not product code.
//
// The 14 GSL functions the reference's CUDA solver calls (declared in gsl/gsl_matrix.h of this directory),
// written from GSL's documented semantics: row-major matrices addressed through `tda`, LU with partial
// pivoting (P A = L U, unit lower triangle stored below the diagonal), inverse from the factor.
// The reference pokes size1/size2/tda between calls (src/sypha_solver_init.cpp:570-611), so every routine
// reads them at call time and never assumes tda == size2.
//
// Deliberate difference from the GSL the reference links: GSL's CBLAS is single-threaded; the three big
// products and the inverse here use OpenMP over rows / column slices (same arithmetic per entry, terms
// added in the same k order), because a faithful single-thread version would spend a minute per LP in the
// starting point at m = 1000.  That only shortens the reference's "start" phase; the iterations/s metric is
// taken from its loop timer (sypha_solver.cpp:487,821) and is not affected.
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "gsl/gsl_matrix.h"

extern "C" {

gsl_vector *gsl_vector_alloc(size_t n)
{
    gsl_vector *v = static_cast<gsl_vector *>(std::malloc(sizeof(gsl_vector)));
    v->size = n;
    v->stride = 1;
    v->data = static_cast<double *>(std::malloc(sizeof(double) * std::max<size_t>(n, 1)));
    v->block = nullptr;
    v->owner = 1;
    return v;
}

void gsl_vector_free(gsl_vector *v)
{
    if (!v) return;
    std::free(v->data);
    std::free(v);
}

double gsl_vector_min(const gsl_vector *v)
{
    double m = v->data[0];
    for (size_t i = 1; i < v->size; ++i)
    {
        const double x = v->data[i * v->stride];
        if (x < m) m = x;
        if (std::isnan(x)) return x;
    }
    return m;
}

int gsl_vector_add_constant(gsl_vector *v, double x)
{
    for (size_t i = 0; i < v->size; ++i) v->data[i * v->stride] += x;
    return 0;
}

gsl_matrix *gsl_matrix_calloc(size_t n1, size_t n2)
{
    gsl_matrix *m = static_cast<gsl_matrix *>(std::malloc(sizeof(gsl_matrix)));
    m->size1 = n1;
    m->size2 = n2;
    m->tda = n2;
    m->data = static_cast<double *>(std::calloc(std::max<size_t>(n1 * n2, 1), sizeof(double)));
    m->block = nullptr;
    m->owner = 1;
    return m;
}

void gsl_matrix_free(gsl_matrix *m)
{
    if (!m) return;
    std::free(m->data);
    std::free(m);
}

gsl_permutation *gsl_permutation_alloc(size_t n)
{
    gsl_permutation *p = static_cast<gsl_permutation *>(std::malloc(sizeof(gsl_permutation)));
    p->size = n;
    p->data = static_cast<size_t *>(std::malloc(sizeof(size_t) * std::max<size_t>(n, 1)));
    return p;
}

void gsl_permutation_free(gsl_permutation *p)
{
    if (!p) return;
    std::free(p->data);
    std::free(p);
}

// C = alpha op(A) op(B) + beta C.  Row i of C is owned by one thread; terms are added for k ascending.
int gsl_blas_dgemm(CBLAS_TRANSPOSE_t ta, CBLAS_TRANSPOSE_t tb, double alpha, const gsl_matrix *A,
                   const gsl_matrix *B, double beta, gsl_matrix *C)
{
    const bool at = (ta != CblasNoTrans), bt = (tb != CblasNoTrans);
    const long M = static_cast<long>(C->size1), N = static_cast<long>(C->size2);
    const long K = static_cast<long>(at ? A->size1 : A->size2);
    const long KB = static_cast<long>(bt ? B->size2 : B->size1);
    if (static_cast<long>(at ? A->size2 : A->size1) != M || static_cast<long>(bt ? B->size1 : B->size2) != N || K != KB)
        return 19; /* GSL_EBADLEN */
    const double *a = A->data, *b = B->data;
    double *c = C->data;
    const long lda = static_cast<long>(A->tda), ldb = static_cast<long>(B->tda), ldc = static_cast<long>(C->tda);
#pragma omp parallel for schedule(dynamic, 8)
    for (long i = 0; i < M; ++i)
    {
        double *ci = c + i * ldc;
        if (beta == 0.0)
            for (long j = 0; j < N; ++j) ci[j] = 0.0;
        else if (beta != 1.0)
            for (long j = 0; j < N; ++j) ci[j] *= beta;
        if (bt)
        {
            // dot products of row-like operands: op(B)(k, j) = B[j][k]
            for (long j = 0; j < N; ++j)
            {
                const double *bj = b + j * ldb;
                double acc = 0.0;
                if (!at)
                {
                    const double *ai = a + i * lda;
                    for (long k = 0; k < K; ++k) acc += ai[k] * bj[k];
                }
                else
                    for (long k = 0; k < K; ++k) acc += a[k * lda + i] * bj[k];
                ci[j] += alpha * acc;
            }
        }
        else
        {
            // row updates: C[i][:] += (alpha op(A)(i,k)) * B[k][:]; a zero multiplier adds exact zeros
            for (long k = 0; k < K; ++k)
            {
                const double aik = alpha * (at ? a[k * lda + i] : a[i * lda + k]);
                if (aik == 0.0) continue;
                const double *bk = b + k * ldb;
                for (long j = 0; j < N; ++j) ci[j] += aik * bk[j];
            }
        }
    }
    return 0;
}

int gsl_blas_dgemv(CBLAS_TRANSPOSE_t ta, double alpha, const gsl_matrix *A, const gsl_vector *x, double beta,
                   gsl_vector *y)
{
    const long M = static_cast<long>(A->size1), N = static_cast<long>(A->size2), lda = static_cast<long>(A->tda);
    const double *a = A->data;
    if (ta == CblasNoTrans)
    {
        if (static_cast<long>(x->size) != N || static_cast<long>(y->size) != M) return 19;
#pragma omp parallel for schedule(static)
        for (long i = 0; i < M; ++i)
        {
            const double *ai = a + i * lda;
            double acc = 0.0;
            for (long j = 0; j < N; ++j) acc += ai[j] * x->data[j * x->stride];
            double &yi = y->data[i * y->stride];
            yi = alpha * acc + (beta == 0.0 ? 0.0 : beta * yi);
        }
    }
    else
    {
        if (static_cast<long>(x->size) != M || static_cast<long>(y->size) != N) return 19;
        for (long j = 0; j < N; ++j)
        {
            double &yj = y->data[j * y->stride];
            yj = (beta == 0.0 ? 0.0 : beta * yj);
        }
        for (long i = 0; i < M; ++i)
        {
            const double t = alpha * x->data[i * x->stride];
            if (t == 0.0) continue;
            const double *ai = a + i * lda;
            for (long j = 0; j < N; ++j) y->data[j * y->stride] += t * ai[j];
        }
    }
    return 0;
}

int gsl_blas_ddot(const gsl_vector *x, const gsl_vector *y, double *result)
{
    double acc = 0.0;
    for (size_t i = 0; i < x->size; ++i) acc += x->data[i * x->stride] * y->data[i * y->stride];
    *result = acc;
    return 0;
}

// P A = L U in place, partial pivoting (largest magnitude in the column), p->data[i] = source row of row i.
int gsl_linalg_LU_decomp(gsl_matrix *A, gsl_permutation *p, int *signum)
{
    const long n = static_cast<long>(A->size1), lda = static_cast<long>(A->tda);
    if (A->size1 != A->size2 || p->size != A->size1) return 19;
    double *a = A->data;
    *signum = 1;
    for (long i = 0; i < n; ++i) p->data[i] = static_cast<size_t>(i);
    for (long j = 0; j + 1 < n; ++j)
    {
        long piv = j;
        double best = std::fabs(a[j * lda + j]);
        for (long i = j + 1; i < n; ++i)
        {
            const double v = std::fabs(a[i * lda + j]);
            if (v > best) { best = v; piv = i; }
        }
        if (piv != j)
        {
            double *r0 = a + j * lda, *r1 = a + piv * lda;
            for (long k = 0; k < n; ++k) std::swap(r0[k], r1[k]);
            std::swap(p->data[j], p->data[piv]);
            *signum = -*signum;
        }
        const double ajj = a[j * lda + j];
        if (ajj != 0.0)
        {
            const double *rj = a + j * lda;
#pragma omp parallel for schedule(static) if (n - j > 256)
            for (long i = j + 1; i < n; ++i)
            {
                double *ri = a + i * lda;
                const double l = ri[j] / ajj;
                ri[j] = l;
                if (l != 0.0)
                    for (long k = j + 1; k < n; ++k) ri[k] -= l * rj[k];
            }
        }
    }
    return 0;
}

// inverse = U^-1 L^-1 P: each thread sweeps its own slice of columns of the right-hand side P.
int gsl_linalg_LU_invert(const gsl_matrix *LU, const gsl_permutation *p, gsl_matrix *inverse)
{
    const long n = static_cast<long>(LU->size1), lda = static_cast<long>(LU->tda), ldi = static_cast<long>(inverse->tda);
    if (LU->size1 != LU->size2 || inverse->size1 != LU->size1 || inverse->size2 != LU->size1) return 19;
    const double *a = LU->data;
    double *x = inverse->data;
    for (long i = 0; i < n; ++i)
        if (a[i * lda + i] == 0.0) return 1; /* GSL_EDOM: singular */
    for (long i = 0; i < n; ++i)
    {
        double *xi = x + i * ldi;
        for (long k = 0; k < n; ++k) xi[k] = 0.0;
        xi[p->data[i]] = 1.0;
    }
    const long slice = 256;
#pragma omp parallel for schedule(dynamic, 1)
    for (long c0 = 0; c0 < n; c0 += slice)
    {
        const long c1 = std::min(n, c0 + slice);
        for (long i = 1; i < n; ++i)            // L y = P e  (unit diagonal)
        {
            double *xi = x + i * ldi;
            const double *li = a + i * lda;
            for (long k = 0; k < i; ++k)
            {
                const double l = li[k];
                if (l == 0.0) continue;
                const double *xk = x + k * ldi;
                for (long c = c0; c < c1; ++c) xi[c] -= l * xk[c];
            }
        }
        for (long i = n - 1; i >= 0; --i)       // U x = y
        {
            double *xi = x + i * ldi;
            const double *ui = a + i * lda;
            for (long k = i + 1; k < n; ++k)
            {
                const double u = ui[k];
                if (u == 0.0) continue;
                const double *xk = x + k * ldi;
                for (long c = c0; c < c1; ++c) xi[c] -= u * xk[c];
            }
            const double d = ui[i];
            for (long c = c0; c < c1; ++c) xi[c] /= d;
        }
    }
    return 0;
}

} // extern "C"
