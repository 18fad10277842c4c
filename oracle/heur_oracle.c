/* oracle/heur_oracle.c - TEST INFRASTRUCTURE (CPU restatement; never linked into the product).
 *
 * The per-node combinatorial rules of the reference's branch-and-bound driver, restated in plain C from
 * /root/reference/src/sypha_solver_heuristics.cpp and /root/reference/src/sypha_solver_bnb.cpp:
 *
 *   oracle_select_branch ............ collect_fractional_candidates (bnb.cpp:368-382) +
 *                                     MostFractionalSelector (heuristics.cpp:10-30) /
 *                                     HighestCostFractionalSelector (:32-51)
 *   oracle_nearest_integer_fixing ... NearestIntegerFixingHeuristic::tryBuild (:53-110)
 *   oracle_dual_guided_cover_repair . DualGuidedCoverRepairHeuristic::tryBuild (:112-292)
 *
 * Same decisions, same floating-point sums in the same order (uncovered rows ascending), same first-index
 * tie-breaks (`>` / `<` scans).  The reference finds "the entry of column j in row i" by scanning every
 * uncovered row for every column (O(rounds * n * nnz)); this restatement walks the column's own row list
 * (rows ascending), which visits the same entries in the same order.  One freedom the reference leaves open:
 * the redundancy pass sorts the selected columns by cost with std::sort (:248-250), whose order among EQUAL
 * costs is unspecified; here equal costs go by ascending column index.
 *
 * Pinned against the reference's own translation unit compiled as it lies (oracle/Makefile ->
 * oracle/_ref/libref_heur.so) by tests/test_heuristics_oracle.py.
 *
 * Model: base rows in CSR (m rows; columns >= n0 are slack/surplus columns and are skipped, as the reference's
 * `col < ncolsOriginal` tests do), obj[n0], rhs[m].  Decisions: (var, fix) pairs.  Build:
 *   gcc -O2 -shared -fPIC -ffp-contract=off oracle/heur_oracle.c -o oracle/_build/libheur_oracle.so
 */
#include <float.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

typedef struct
{
    int *ptr, *rows;
    double *vals;
} csc_t;

/* column lists of the original columns, rows ascending, positive and non-positive entries alike */
static csc_t build_csc(int m, int n0, const int *offs, const int *inds, const double *vals)
{
    csc_t c;
    c.ptr = (int *)calloc((size_t)n0 + 1, sizeof(int));
    for (int i = 0; i < m; ++i)
        for (int k = offs[i]; k < offs[i + 1]; ++k)
            if (inds[k] >= 0 && inds[k] < n0) c.ptr[inds[k] + 1]++;
    for (int j = 0; j < n0; ++j) c.ptr[j + 1] += c.ptr[j];
    c.rows = (int *)malloc(sizeof(int) * (size_t)(c.ptr[n0] > 0 ? c.ptr[n0] : 1));
    c.vals = (double *)malloc(sizeof(double) * (size_t)(c.ptr[n0] > 0 ? c.ptr[n0] : 1));
    int *fill = (int *)calloc((size_t)n0 > 0 ? (size_t)n0 : 1, sizeof(int));
    for (int i = 0; i < m; ++i)
        for (int k = offs[i]; k < offs[i + 1]; ++k)
        {
            const int j = inds[k];
            if (j < 0 || j >= n0) continue;
            const int p = c.ptr[j] + fill[j]++;
            c.rows[p] = i;
            c.vals[p] = vals[k];
        }
    free(fill);
    return c;
}

static void free_csc(csc_t *c)
{
    free(c->ptr);
    free(c->rows);
    free(c->vals);
}

/* rule 0: most fractional, rule 1: highest cost among the fractional candidates.  Returns the column or -1
 * (no candidate: the point is integral, is_binary_integral_solution bnb.cpp:350-366); *frac_out = its
 * |x - floor(x + 0.5)|. */
int oracle_select_branch(int rule, const double *x, const double *obj, int n0, double tol, double *frac_out)
{
    int best = -1;
    double best_score = rule == 0 ? -1.0 : -INFINITY;
    for (int j = 0; j < n0; ++j)
    {
        const double v = x[j], nearest = floor(v + 0.5);
        if (!((fabs(v - nearest) > tol) || (nearest < -tol) || (nearest > 1.0 + tol))) continue;
        const double score = rule == 0 ? fabs(v - nearest) : obj[j];
        if (score > best_score)
        {
            best_score = score;
            best = j;
        }
    }
    if (frac_out) *frac_out = best >= 0 ? fabs(x[best] - floor(x[best] + 0.5)) : 0.0;
    return best;
}

int oracle_nearest_integer_fixing(int m, int n0, const int *offs, const int *inds, const double *vals,
                                  const double *obj, const double *rhs, const double *x, int ndec, const int *dvar,
                                  const int *dfix, double tol, double *sol, double *objective)
{
    for (int j = 0; j < n0; ++j)
    {
        const double r = floor(x[j] + 0.5);
        sol[j] = r < 0.0 ? 0.0 : (r > 1.0 ? 1.0 : r);
    }
    for (int d = 0; d < ndec; ++d)
        if (dvar[d] >= 0 && dvar[d] < n0) sol[dvar[d]] = (double)dfix[d];
    *objective = INFINITY;
    for (int i = 0; i < m; ++i)
    {
        double coverage = 0.0;
        for (int k = offs[i]; k < offs[i + 1]; ++k)
            if (inds[k] >= 0 && inds[k] < n0) coverage += vals[k] * sol[inds[k]];
        if (coverage + tol < rhs[i]) return 0;
    }
    double o = 0.0;
    for (int j = 0; j < n0; ++j) o += obj[j] * sol[j];
    *objective = o;
    return 1;
}

static void recompute_coverage(int m, int n0, const int *offs, const int *inds, const double *vals, const double *sol,
                               double *coverage)
{
    for (int i = 0; i < m; ++i)
    {
        double c = 0.0;
        for (int k = offs[i]; k < offs[i + 1]; ++k)
            if (inds[k] >= 0 && inds[k] < n0 && sol[inds[k]] > 0.5) c += vals[k];
        coverage[i] = c;
    }
}

typedef struct
{
    double cost;
    int col;
} sel_t;

static int by_cost_desc(const void *a, const void *b)
{
    const sel_t *p = (const sel_t *)a, *q = (const sel_t *)b;
    if (p->cost != q->cost) return p->cost > q->cost ? -1 : 1;
    return p->col < q->col ? -1 : (p->col > q->col ? 1 : 0);
}

/* ny = length of y (the reference tests relaxedDual.size() > i).  *steps_out = columns added by the repair. */
int oracle_dual_guided_cover_repair(int m, int n0, const int *offs, const int *inds, const double *vals,
                                    const double *obj, const double *rhs, const double *x, const double *y, int ny,
                                    int ndec, const int *dvar, const int *dfix, double tol, double *sol,
                                    double *objective, int *steps_out)
{
    char *fixed0 = (char *)calloc((size_t)n0 + 1, 1), *fixed1 = (char *)calloc((size_t)n0 + 1, 1);
    double *coverage = (double *)malloc(sizeof(double) * (size_t)(m > 0 ? m : 1));
    csc_t c = build_csc(m, n0, offs, inds, vals);
    int steps = 0, feasible = 0;
    *objective = INFINITY;
    memset(sol, 0, sizeof(double) * (size_t)n0);
    for (int d = 0; d < ndec; ++d)
    {
        if (dvar[d] < 0 || dvar[d] >= n0) continue;
        if (dfix[d] == 0) fixed0[dvar[d]] = 1;
        else
        {
            fixed1[dvar[d]] = 1;
            sol[dvar[d]] = 1.0;
        }
    }
    for (int j = 0; j < n0; ++j)
    {
        if (fixed0[j]) { sol[j] = 0.0; continue; }
        if (fixed1[j]) continue;
        if (x[j] >= 1.0 - tol) sol[j] = 1.0;
    }
    recompute_coverage(m, n0, offs, inds, vals, sol, coverage);
#define COVERED(i) (coverage[i] + tol >= rhs[i])
    for (;;)
    {
        int uncovered = -1;
        for (int i = 0; i < m; ++i)
            if (!COVERED(i)) { uncovered = i; break; }
        if (uncovered < 0) break;
        int best_col = -1;
        double best_score = -INFINITY;
        for (int j = 0; j < n0; ++j)
        {
            if (sol[j] > 0.5 || fixed0[j]) continue;
            double ug = 0.0, dg = 0.0;
            for (int p = c.ptr[j]; p < c.ptr[j + 1]; ++p)
            {
                const int i = c.rows[p];
                if (COVERED(i)) continue;
                const double aij = c.vals[p];
                if (aij > 0.0)
                {
                    ug += aij;
                    if (ny > i) dg += (y[i] > 0.0 ? y[i] : 0.0) * aij;
                }
            }
            if (ug <= 0.0) continue;
            const double cost = obj[j] > 1e-9 ? obj[j] : 1e-9;
            const double score = (ug + dg) / cost;
            if (score > best_score)
            {
                best_score = score;
                best_col = j;
            }
        }
        if (best_col < 0)
        {   /* fallback (:215-243): cheapest usable column with a positive entry in an uncovered row */
            double best_cost = INFINITY;
            for (int i = 0; i < m; ++i)
            {
                if (COVERED(i)) continue;
                for (int k = offs[i]; k < offs[i + 1]; ++k)
                {
                    const int col = inds[k];
                    if (col < 0 || col >= n0 || fixed0[col] || sol[col] > 0.5) continue;
                    if (vals[k] <= 0.0) continue;
                    if (obj[col] < best_cost)
                    {
                        best_cost = obj[col];
                        best_col = col;
                    }
                }
            }
            if (best_col < 0) goto done;
        }
        sol[best_col] = 1.0;
        ++steps;
        recompute_coverage(m, n0, offs, inds, vals, sol, coverage);
    }
    {
        sel_t *sel = (sel_t *)malloc(sizeof(sel_t) * (size_t)(n0 > 0 ? n0 : 1));
        int ns = 0;
        for (int j = 0; j < n0; ++j)
            if (sol[j] > 0.5 && !fixed1[j])
            {
                sel[ns].cost = obj[j];
                sel[ns++].col = j;
            }
        qsort(sel, (size_t)ns, sizeof(sel_t), by_cost_desc);
        for (int t = 0; t < ns; ++t)
        {
            sol[sel[t].col] = 0.0;
            recompute_coverage(m, n0, offs, inds, vals, sol, coverage);
            int ok = 1;
            for (int i = 0; i < m; ++i)
                if (!COVERED(i)) { ok = 0; break; }
            if (!ok)
            {
                sol[sel[t].col] = 1.0;
                recompute_coverage(m, n0, offs, inds, vals, sol, coverage);
            }
        }
        free(sel);
    }
    feasible = 1;
    for (int i = 0; i < m; ++i)
        if (!COVERED(i)) { feasible = 0; break; }
    if (feasible)
    {
        double o = 0.0;
        for (int j = 0; j < n0; ++j) o += obj[j] * sol[j];
        *objective = o;
    }
done:
#undef COVERED
    if (steps_out) *steps_out = steps;
    free_csc(&c);
    free(fixed0);
    free(fixed1);
    free(coverage);
    return feasible;
}
