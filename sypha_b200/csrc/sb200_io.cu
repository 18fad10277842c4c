// sb200_io.cu - host side of the path's input: the OR-Library set-covering text format and the standard form
// the solver consumes, in one pass over the file.
//
// Restates /root/reference/src/model_reader.cpp:90-174 (model_reader_read_scp_file_sparse_csr): tokens are
// `m n`, n objective coefficients, then per row `k idx_1 .. idx_k` (1-based); the standard form is
// A = [A0 | -I] (the surplus entry AFTER the row's own entries, :146-147), b = 1, c = [c0; 0].  The reference
// scans with fscanf (one libc call per token) and grows three std::vectors by push_back; through its public API
// the model is then rebuilt entry by entry (Constraint::SetCoefficient is O(row length) per call,
// src/sypha_api.cpp:41-49 - SURVEY.md 8f rank 3).  Here the file is read once, tokens are parsed in place and
// the CSR arrays are written directly at their final size.  Host code only; no CUDA call.
#include "../../include/sypha_b200.h"

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

namespace {

struct Scanner
{
    const char *p, *end;
    void skip()
    {
        while (p < end && (*p == ' ' || *p == '\n' || *p == '\r' || *p == '\t' || *p == '\f' || *p == '\v')) ++p;
    }
    bool next_int(long long &v)
    {
        skip();
        if (p >= end) return false;
        bool neg = false;
        if (*p == '-' || *p == '+') neg = (*p++ == '-');
        if (p >= end || *p < '0' || *p > '9') return false;
        long long x = 0;
        while (p < end && *p >= '0' && *p <= '9') x = x * 10 + (*p++ - '0');
        v = neg ? -x : x;
        return true;
    }
    bool next_double(double &v)
    {
        skip();
        if (p >= end) return false;
        // plain integers (every OR-Library cost) take the fast path; anything else goes through strtod
        const char *q = p;
        long long x;
        if (next_int(x) && (p >= end || (*p != '.' && *p != 'e' && *p != 'E')))
        {
            v = (double)x;
            return true;
        }
        p = q;
        char *stop = nullptr;
        v = strtod(p, &stop);
        if (stop == p) return false;
        p = stop;
        return true;
    }
};

} // namespace

extern "C" int sb200_read_scp(const char *path, sb200_scp_model *out)
{
    if (!path || !out) return SB200_ERR_INVALID;
    memset(out, 0, sizeof *out);
    FILE *f = fopen(path, "rb");
    if (!f) return SB200_ERR_INVALID;
    fseek(f, 0, SEEK_END);
    const long size = ftell(f);
    fseek(f, 0, SEEK_SET);
    if (size <= 0) { fclose(f); return SB200_ERR_INVALID; }
    std::vector<char> buf((size_t)size + 1);        // + a terminating NUL: strtod on a trailing non-integer token must stop there
    buf[(size_t)size] = '\0';
    const size_t got = fread(buf.data(), 1, (size_t)size, f);
    fclose(f);
    if (got != (size_t)size) return SB200_ERR_INVALID;
    Scanner s{buf.data(), buf.data() + size};

    long long m = 0, n0 = 0;
    if (!s.next_int(m) || !s.next_int(n0) || m <= 0 || n0 <= 0 || m + n0 > 0x7fffffffll) return SB200_ERR_INVALID;
    double *c = (double *)calloc((size_t)(n0 + m), sizeof(double));
    double *b = (double *)malloc(sizeof(double) * (size_t)m);
    int *offs = (int *)malloc(sizeof(int) * (size_t)(m + 1));
    if (!c || !b || !offs) { free(c); free(b); free(offs); return SB200_ERR_NOMEM; }
    auto bail = [&](int code, int *inds, double *vals) {
        free(c); free(b); free(offs); free(inds); free(vals);
        return code;
    };
    for (long long j = 0; j < n0; ++j)
        if (!s.next_double(c[j])) return bail(SB200_ERR_INVALID, nullptr, nullptr);
    // the rows are not announced in advance: grow geometrically (the text is at least 2 bytes per entry)
    size_t cap = (size_t)size / 2 + (size_t)m + 16;
    int *inds = (int *)malloc(sizeof(int) * cap);
    if (!inds) return bail(SB200_ERR_NOMEM, nullptr, nullptr);
    size_t nnz = 0;
    offs[0] = 0;
    for (long long i = 0; i < m; ++i)
    {
        long long k = 0;
        if (!s.next_int(k) || k < 0) return bail(SB200_ERR_INVALID, inds, nullptr);
        if (nnz + (size_t)k + 1 > cap) return bail(SB200_ERR_INVALID, inds, nullptr);   // more entries than bytes: malformed
        for (long long q = 0; q < k; ++q)
        {
            long long idx = 0;
            if (!s.next_int(idx) || idx < 1 || idx > n0) return bail(SB200_ERR_INVALID, inds, nullptr);
            inds[nnz++] = (int)(idx - 1);
        }
        inds[nnz++] = (int)(n0 + i);                 // surplus column, last in its row
        if (nnz > 0x7fffffffull) return bail(SB200_ERR_UNSUPPORTED, inds, nullptr);
        offs[i + 1] = (int)nnz;
        b[i] = 1.0;
    }
    double *vals = (double *)malloc(sizeof(double) * (nnz ? nnz : 1));
    if (!vals) return bail(SB200_ERR_NOMEM, inds, nullptr);
    for (size_t e = 0; e < nnz; ++e) vals[e] = 1.0;
    for (long long i = 0; i < m; ++i) vals[offs[i + 1] - 1] = -1.0;
    out->m = (int)m;
    out->n = (int)(n0 + m);
    out->n_orig = (int)n0;
    out->nnz = (long long)nnz;
    out->csr_offs = offs;
    out->csr_inds = inds;
    out->csr_vals = vals;
    out->c = c;
    out->b = b;
    return SB200_OK;
}

extern "C" void sb200_free_scp(sb200_scp_model *mdl)
{
    if (!mdl) return;
    free(mdl->csr_offs);
    free(mdl->csr_inds);
    free(mdl->csr_vals);
    free(mdl->c);
    free(mdl->b);
    memset(mdl, 0, sizeof *mdl);
}

// ---- general row model -> standard form (sypha::Solver::Impl::buildStandardForm, src/sypha_api.cpp:136-250) ---------------
namespace {
enum RowKind { ROW_EQ, ROW_GE, ROW_LE, ROW_RANGE, ROW_FREE };
inline RowKind row_kind(double lb, double ub)
{
    const bool has_lb = std::isfinite(lb), has_ub = std::isfinite(ub);
    if (has_lb && has_ub) return std::fabs(lb - ub) <= 1e-15 ? ROW_EQ : ROW_RANGE;     // sypha_api.cpp:145,164
    if (has_lb) return ROW_GE;
    if (has_ub) return ROW_LE;
    return ROW_FREE;
}
inline bool row_model_ok(const sb200_row_model *in)
{
    if (!in || in->n_vars < 0 || in->n_rows < 0 || !in->row_offs || (in->n_rows && (!in->row_lb || !in->row_ub))) return false;
    if (in->row_offs[0] != 0) return false;
    for (int i = 0; i < in->n_rows; ++i)
        if (in->row_offs[i + 1] < in->row_offs[i]) return false;
    if (in->row_offs[in->n_rows] > 0 && (!in->row_inds || !in->row_vals)) return false;
    return true;
}
} // namespace

extern "C" int sb200_standard_form_size(const sb200_row_model *in, int *nrows, int *ncols, long long *nnz)
{
    if (!row_model_ok(in) || !nrows || !ncols || !nnz) return SB200_ERR_INVALID;
    long long rows = 0, slacks = 0, entries = 0;
    for (int i = 0; i < in->n_rows; ++i)
    {
        const long long len = in->row_offs[i + 1] - in->row_offs[i];
        switch (row_kind(in->row_lb[i], in->row_ub[i]))
        {
        case ROW_EQ: case ROW_FREE: rows += 1; entries += len; break;
        case ROW_GE: case ROW_LE: rows += 1; slacks += 1; entries += len + 1; break;
        case ROW_RANGE: rows += 2; slacks += 2; entries += 2 * (len + 1); break;
        }
    }
    if (rows > 0x7fffffffll || in->n_vars + slacks > 0x7fffffffll || entries > 0x7fffffffll) return SB200_ERR_UNSUPPORTED;
    *nrows = (int)rows;
    *ncols = (int)(in->n_vars + slacks);
    *nnz = entries;
    return SB200_OK;
}

extern "C" int sb200_build_standard_form(const sb200_row_model *in, int *csr_offs, int *csr_inds, double *csr_vals, double *obj,
                                         double *rhs)
{
    int nrows = 0, ncols = 0;
    long long nnz = 0;
    const int rc = sb200_standard_form_size(in, &nrows, &ncols, &nnz);
    if (rc != SB200_OK) return rc;
    if (!csr_offs || !obj || (nrows && !rhs) || (nnz && (!csr_inds || !csr_vals))) return SB200_ERR_INVALID;
    const int n = in->n_vars;
    for (long long t = 0; t < in->row_offs[in->n_rows]; ++t)
        if (in->row_inds[t] < 0 || in->row_inds[t] >= n) return SB200_ERR_INVALID;
    for (int j = 0; j < ncols; ++j) obj[j] = 0.0;                                 // surplus columns cost nothing
    if (in->obj)
        for (int j = 0; j < n; ++j) obj[j] = in->maximize ? -in->obj[j] : in->obj[j];   // sypha_api.cpp:189-196
    int r = 0, slack = n;
    long long o = 0;
    csr_offs[0] = 0;
    // one output row: the coefficients (negated for a "<= ub" row), then the surplus entry; sypha_api.cpp:206-246
    auto emit = [&](int i, bool negate, bool with_slack, double rhs_val) {
        for (int t = in->row_offs[i]; t < in->row_offs[i + 1]; ++t)
        {
            csr_inds[o] = in->row_inds[t];
            csr_vals[o] = negate ? -in->row_vals[t] : in->row_vals[t];
            ++o;
        }
        if (with_slack)
        {
            csr_inds[o] = slack++;
            csr_vals[o] = -1.0;
            ++o;
        }
        rhs[r] = rhs_val;
        csr_offs[++r] = (int)o;
    };
    for (int i = 0; i < in->n_rows; ++i)
    {
        const double lb = in->row_lb[i], ub = in->row_ub[i];
        switch (row_kind(lb, ub))
        {
        case ROW_EQ: emit(i, false, false, lb); break;
        case ROW_FREE: emit(i, false, false, 0.0); break;
        case ROW_GE: emit(i, false, true, lb); break;
        case ROW_LE: emit(i, true, true, -ub); break;
        case ROW_RANGE:
            emit(i, false, true, lb);
            emit(i, true, true, -ub);
            break;
        }
    }
    return SB200_OK;
}
