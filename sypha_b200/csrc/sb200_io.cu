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
