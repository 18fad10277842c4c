// sb200_assemble.cu - normal-equations assembly M = A diag(d) A'.
//
// The reference never forms M: it factors the full (2n+m)^2 KKT matrix
// (/root/reference/src/sypha_solver.cpp:84-92,113-186) or, on the Krylov branch, applies
// A D A' matrix-free (/root/reference/src/sypha_solver_krylov.cu:303-329).  north_star item (1)
// replaces both with an explicit symmetric product.
//
// Sparse path (symbolic once per model, numeric once per IPM iteration):
//   M[i][k] = sum over j in row_i ∩ row_k of a_ij a_kj d_j.
//   Symbolic: for every column j emit the pairs (i >= k) of its rows keyed by the packed
//   lower-triangular index, stable radix sort by key -> per-entry lists of column ids (sorted by j,
//   so the summation order is fixed -> deterministic), prefix sum of the per-entry counts.
//   Numeric: one thread per matrix entry gathers d over its list: no atomics, no symbolic work,
//   coalesced writes of M.  Algorithmic bytes per iteration: 4 B (8+4 B when A has non-unit
//   products) per term + 4 B per entry pointer + 8 B per written entry.
// Dense path: FP64 tensor-core SYRK C = A diag(d) A' on a dense copy of A (64x64 tiles, DMMA).
#include "sb200_kernels.cuh"
#include <algorithm>
#include "sb200_dmma.cuh"

#include <cub/cub.cuh>

namespace sb200 {

// ---------------------------------------------------------------------------------------------
// CSR -> CSC (device, deterministic: stable sort keeps rows ascending inside each column)
// ---------------------------------------------------------------------------------------------
__global__ void k_expand_rows(int m, const int *__restrict__ offs, int *__restrict__ rowid,
                              unsigned int *__restrict__ pos)
{
    const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
    for (int row = blockIdx.x * wpb + (threadIdx.x >> 5); row < m; row += gridDim.x * wpb)
        for (int k = offs[row] + lane; k < offs[row + 1]; k += 32)
        {
            rowid[k] = row;
            pos[k] = (unsigned int)k;
        }
}
__global__ void k_count_cols(long long nnz, const int *__restrict__ inds, int *__restrict__ cnt)
{
    for (long long k = blockIdx.x * (long long)blockDim.x + threadIdx.x; k < nnz;
         k += (long long)gridDim.x * blockDim.x)
        atomicAdd(&cnt[inds[k]], 1);
}
__global__ void k_gather_csc(long long nnz, const unsigned int *__restrict__ perm,
                             const int *__restrict__ rowid, const double *__restrict__ vals,
                             int *__restrict__ out_rows, double *__restrict__ out_vals)
{
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < nnz;
         t += (long long)gridDim.x * blockDim.x)
    {
        const unsigned int p = perm[t];
        out_rows[t] = rowid[p];
        out_vals[t] = vals[p];
    }
}

static int bits_for(unsigned long long n)
{
    int b = 1;
    while (b < 64 && (1ull << b) < n) ++b;
    return b;
}

int build_csc(ErrorSink &err, int m, int n, long long nnz, const int *csr_offs, const int *csr_inds,
              const double *csr_vals, int *csc_colptr, int *csc_rows, double *csc_vals,
              cudaStream_t st)
{
    int *rowid = nullptr, *cnt = nullptr;
    unsigned int *pos = nullptr, *perm = nullptr;
    int *keys_out = nullptr;
    void *tmp = nullptr;
    size_t tmp_bytes = 0, scan_bytes = 0;
    SB200_CUDA_TRY(err, cudaMallocAsync(&rowid, sizeof(int) * (size_t)nnz, st));
    SB200_CUDA_TRY(err, cudaMallocAsync(&pos, sizeof(int) * (size_t)nnz, st));
    SB200_CUDA_TRY(err, cudaMallocAsync(&perm, sizeof(int) * (size_t)nnz, st));
    SB200_CUDA_TRY(err, cudaMallocAsync(&keys_out, sizeof(int) * (size_t)nnz, st));
    SB200_CUDA_TRY(err, cudaMallocAsync(&cnt, sizeof(int) * (size_t)(n + 1), st));
    SB200_CUDA_TRY(err, cudaMemsetAsync(cnt, 0, sizeof(int) * (size_t)(n + 1), st));
    k_expand_rows<<<grid_for((long long)m * 32, 256, 148 * 16), 256, 0, st>>>(m, csr_offs, rowid, pos);
    k_count_cols<<<grid_for(nnz, 256, 148 * 16), 256, 0, st>>>(nnz, csr_inds, cnt);
    const int nb = bits_for((unsigned long long)n);
    SB200_CUDA_TRY(err, cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, csr_inds, keys_out, pos, perm,
                                                        nnz, 0, nb, st));
    SB200_CUDA_TRY(err, cub::DeviceScan::ExclusiveSum(nullptr, scan_bytes, cnt, csc_colptr, n + 1, st));
    if (scan_bytes > tmp_bytes) tmp_bytes = scan_bytes;
    SB200_CUDA_TRY(err, cudaMallocAsync(&tmp, tmp_bytes, st));
    size_t tb = tmp_bytes;
    SB200_CUDA_TRY(err, cub::DeviceRadixSort::SortPairs(tmp, tb, csr_inds, keys_out, pos, perm, nnz, 0, nb, st));
    tb = tmp_bytes;
    SB200_CUDA_TRY(err, cub::DeviceScan::ExclusiveSum(tmp, tb, cnt, csc_colptr, n + 1, st));
    k_gather_csc<<<grid_for(nnz, 256, 148 * 16), 256, 0, st>>>(nnz, perm, rowid, csr_vals, csc_rows, csc_vals);
    g_launch_count += 3;
    SB200_CUDA_TRY(err, cudaStreamSynchronize(st));
    cudaFreeAsync(rowid, st); cudaFreeAsync(pos, st); cudaFreeAsync(perm, st); cudaFreeAsync(keys_out, st); cudaFreeAsync(cnt, st); cudaFreeAsync(tmp, st);
    return SB200_OK;
}

// ---------------------------------------------------------------------------------------------
// symbolic structure of M
// ---------------------------------------------------------------------------------------------
__global__ void k_col_term_counts(int n, const int *__restrict__ colptr, const double *__restrict__ vals,
                                  unsigned long long *__restrict__ cnt, int *__restrict__ general)
{
    for (int j = blockIdx.x * blockDim.x + threadIdx.x; j <= n; j += gridDim.x * blockDim.x)
    {
        if (j == n)
        {
            cnt[j] = 0;
            continue;
        }
        const int a = colptr[j], e = colptr[j + 1];
        const unsigned long long c = (unsigned long long)(e - a);
        cnt[j] = c * (c + 1) / 2;
        // all products a_ij a_kj are +1 iff the column is all +1 or all -1
        bool unit = true;
        if (e > a)
        {
            const double v0 = vals[a];
            unit = (v0 == 1.0 || v0 == -1.0);
            for (int k = a + 1; k < e && unit; ++k)
                unit = (vals[k] == v0);
        }
        if (!unit) atomicOr(general, 1);
    }
}

template <bool GENERAL>
__global__ void k_emit_terms(int n, const int *__restrict__ colptr, const int *__restrict__ rows,
                             const double *__restrict__ vals, const unsigned long long *__restrict__ toff,
                             unsigned int *__restrict__ keys, unsigned int *__restrict__ payload,
                             unsigned int *__restrict__ colj, double *__restrict__ w,
                             unsigned int *__restrict__ pair_cnt)
{
    const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
    for (int j = blockIdx.x * wpb + (threadIdx.x >> 5); j < n; j += gridDim.x * wpb)
    {
        const int a0 = colptr[j];
        const long long c = colptr[j + 1] - a0;
        const long long cnt = c * (c + 1) / 2;
        const unsigned long long base = toff[j];
        for (long long q = lane; q < cnt; q += 32)
        {
            long long a = (long long)((sqrt(8.0 * (double)q + 1.0) - 1.0) * 0.5);
            while ((a + 1) * (a + 2) / 2 <= q) ++a;
            while (a * (a + 1) / 2 > q) --a;
            const long long b = q - a * (a + 1) / 2;
            const unsigned long long ra = (unsigned long long)rows[a0 + a];   // ra >= rb (rows ascending)
            const unsigned long long rb = (unsigned long long)rows[a0 + b];
            const unsigned int key = (unsigned int)(ra * (ra + 1) / 2 + rb);
            keys[base + q] = key;
            if (GENERAL)
            {
                payload[base + q] = (unsigned int)(base + q);
                colj[base + q] = (unsigned int)j;
                w[base + q] = vals[a0 + a] * vals[a0 + b];
            }
            else
                payload[base + q] = (unsigned int)j;
            atomicAdd(&pair_cnt[key], 1u);
        }
    }
}
__global__ void k_gather_terms(long long T, const unsigned int *__restrict__ perm,
                               const unsigned int *__restrict__ colj, const double *__restrict__ w,
                               unsigned int *__restrict__ term_col, double *__restrict__ term_w)
{
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < T;
         t += (long long)gridDim.x * blockDim.x)
    {
        const unsigned int p = perm[t];
        term_col[t] = colj[p];
        term_w[t] = w[p];
    }
}

// ---- compact form of the unit-product case: 2-byte column ids in whole 16-byte chunks -----------------
__global__ void k_pair_chunk_counts(long long n_pairs, const unsigned int *__restrict__ pair_ptr,
                                    unsigned int *__restrict__ cnt)
{
    for (long long p = blockIdx.x * (long long)blockDim.x + threadIdx.x; p <= n_pairs;
         p += (long long)gridDim.x * blockDim.x)
        cnt[p] = p < n_pairs ? (pair_ptr[p + 1] - pair_ptr[p] + 7u) >> 3 : 0u;
}
__global__ void k_pack_terms16(long long n_pairs, const unsigned int *__restrict__ pair_ptr,
                               const unsigned int *__restrict__ term_col, const unsigned int *__restrict__ chunk_ptr,
                               unsigned short *__restrict__ term16, unsigned short pad)
{
    for (long long p = blockIdx.x * (long long)blockDim.x + threadIdx.x; p < n_pairs;
         p += (long long)gridDim.x * blockDim.x)
    {
        const unsigned a = pair_ptr[p], e = pair_ptr[p + 1];
        size_t o = (size_t)chunk_ptr[p] << 3;
        const size_t end = (size_t)chunk_ptr[p + 1] << 3;
        for (unsigned t = a; t < e; ++t)
            term16[o++] = (unsigned short)term_col[t];
        while (o < end)
            term16[o++] = pad;                  // d[n] = 0 by construction of the workspace
    }
}

// The one-block solver sums 16 consecutive entries of a row side by side (a half-warp), chunk position by chunk position,
// and every id is an 8-byte gather from shared memory: 16 random ids hit the 16 double-wide banks three deep on average,
// and that bank-conflict replay is what bounds its assembly.  The order of the ids INSIDE a chunk is free, so it is chosen
// here, once per model: for every group of 16 entries and every chunk position the ids are dealt to the 8 slots so that
// the 16 lanes' ids of one slot fall into different banks as far as possible (greedy: every id goes to the free slot where its bank has been used least; pads last).
__global__ void k_decollide_chunks16(int m, const unsigned int *__restrict__ chunk_ptr, unsigned short *__restrict__ term16,
                                     unsigned short pad)
{
    const int i = blockIdx.y + 1;
    if (i >= m) return;
    const long long p0 = (long long)i * (i + 1) / 2;
    for (int g = blockIdx.x * blockDim.x + threadIdx.x; g * 16 < i; g += gridDim.x * blockDim.x)
    {
        const int k0 = g * 16, cnt = min(16, i - k0);
        unsigned int a[16], len[16];
        unsigned int maxlen = 0;
        for (int l = 0; l < 16; ++l)
        {
            a[l] = l < cnt ? chunk_ptr[p0 + k0 + l] : 0u;
            len[l] = l < cnt ? chunk_ptr[p0 + k0 + l + 1] - a[l] : 0u;
            maxlen = max(maxlen, len[l]);
        }
        for (unsigned int j = 0; j < maxlen; ++j)
        {
            unsigned char load[8][16];                                     // ids per (slot, bank) so far
            for (int q = 0; q < 8; ++q)
                for (int b = 0; b < 16; ++b) load[q][b] = 0;
            for (int l = 0; l < cnt; ++l)
            {
                if (j >= len[l]) continue;
                uint4 *chunk = reinterpret_cast<uint4 *>(term16) + a[l] + j;
                const uint4 v = *chunk;
                const unsigned short ids[8] = {(unsigned short)(v.x & 0xffffu), (unsigned short)(v.x >> 16),
                                               (unsigned short)(v.y & 0xffffu), (unsigned short)(v.y >> 16),
                                               (unsigned short)(v.z & 0xffffu), (unsigned short)(v.z >> 16),
                                               (unsigned short)(v.w & 0xffffu), (unsigned short)(v.w >> 16)};
                unsigned short out[8];
                unsigned int used = 0;
                for (int t = 0; t < 8; ++t)
                {
                    const unsigned short id = ids[t];
                    if (id == pad) continue;
                    const int bank = id & 15;
                    int best = -1, best_load = 1 << 30;
                    for (int qq = 0; qq < 8; ++qq)
                    {   // the free slot where this bank has been used least
                        const int q = (qq + l) & 7;
                        if (!((used >> q) & 1u) && (int)load[q][bank] < best_load)
                        {
                            best = q;
                            best_load = load[q][bank];
                        }
                    }
                    used |= 1u << best;
                    out[best] = id;
                    ++load[best][bank];
                }
                for (int q = 0; q < 8; ++q)
                    if (!((used >> q) & 1u)) out[q] = pad;
                *chunk = make_uint4((unsigned int)out[0] | ((unsigned int)out[1] << 16), (unsigned int)out[2] | ((unsigned int)out[3] << 16),
                                    (unsigned int)out[4] | ((unsigned int)out[5] << 16), (unsigned int)out[6] | ((unsigned int)out[7] << 16));
            }
        }
    }
}

void free_normal_pattern(NormalPattern *p, cudaStream_t st)
{
    if (p->chunk_ptr) cudaFreeAsync(p->chunk_ptr, st);
    if (p->term16) cudaFreeAsync(p->term16, st);
    if (p->pair_ptr) cudaFreeAsync(p->pair_ptr, st);
    if (p->term_col) cudaFreeAsync(p->term_col, st);
    if (p->term_w) cudaFreeAsync(p->term_w, st);
    *p = NormalPattern{};
}

int build_normal_pattern(ErrorSink &err, int m, int n, long long nnz, const int *csc_colptr,
                         const int *csc_rows, const double *csc_vals, NormalPattern *out,
                         cudaStream_t st, int pad_id)
{
    (void)nnz;
    if (pad_id < n) pad_id = n;
    free_normal_pattern(out, st);
    if (m > 65535)
    {
        err.msg = "build_normal_pattern: m > 65535 (packed pair index would overflow 32 bits)";
        return SB200_ERR_UNSUPPORTED;
    }
    const long long n_pairs = (long long)m * (m + 1) / 2;
    unsigned long long *cnt = nullptr, *toff = nullptr;
    int *general_d = nullptr;
    void *tmp = nullptr;
    size_t tmp_bytes = 0;
    SB200_CUDA_TRY(err, cudaMallocAsync(&cnt, sizeof(unsigned long long) * (size_t)(n + 1), st));
    SB200_CUDA_TRY(err, cudaMallocAsync(&toff, sizeof(unsigned long long) * (size_t)(n + 1), st));
    SB200_CUDA_TRY(err, cudaMallocAsync(&general_d, sizeof(int), st));
    SB200_CUDA_TRY(err, cudaMemsetAsync(general_d, 0, sizeof(int), st));
    k_col_term_counts<<<grid_for(n + 1, 256, 148 * 16), 256, 0, st>>>(n, csc_colptr, csc_vals, cnt, general_d);
    SB200_CUDA_TRY(err, cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, cnt, toff, n + 1, st));
    SB200_CUDA_TRY(err, cudaMallocAsync(&tmp, tmp_bytes, st));
    SB200_CUDA_TRY(err, cub::DeviceScan::ExclusiveSum(tmp, tmp_bytes, cnt, toff, n + 1, st));
    unsigned long long T = 0;
    int general = 0;
    SB200_CUDA_TRY(err, cudaMemcpyAsync(&T, toff + n, sizeof T, cudaMemcpyDeviceToHost, st));
    SB200_CUDA_TRY(err, cudaMemcpyAsync(&general, general_d, sizeof general, cudaMemcpyDeviceToHost, st));
    SB200_CUDA_TRY(err, cudaStreamSynchronize(st));
    cudaFreeAsync(tmp, st); tmp = nullptr;
    cudaFreeAsync(cnt, st); cudaFreeAsync(general_d, st);
    if (T >= 0xFFFFFFF0ull)
    {
        cudaFreeAsync(toff, st);
        err.msg = "build_normal_pattern: more than 2^32 product terms; use the PCG strategy";
        return SB200_ERR_UNSUPPORTED;
    }

    unsigned int *keys = nullptr, *keys_out = nullptr, *payload = nullptr, *payload_out = nullptr;
    unsigned int *colj = nullptr, *pair_cnt = nullptr;
    double *w = nullptr;
    const size_t Ts = (size_t)(T > 0 ? T : 1);
    SB200_CUDA_TRY(err, cudaMallocAsync(&keys, 4 * Ts, st));
    SB200_CUDA_TRY(err, cudaMallocAsync(&keys_out, 4 * Ts, st));
    SB200_CUDA_TRY(err, cudaMallocAsync(&payload, 4 * Ts, st));
    SB200_CUDA_TRY(err, cudaMallocAsync(&payload_out, 4 * Ts, st));
    SB200_CUDA_TRY(err, cudaMallocAsync(&pair_cnt, 4 * (size_t)(n_pairs + 1), st));
    SB200_CUDA_TRY(err, cudaMallocAsync(&out->pair_ptr, 4 * (size_t)(n_pairs + 1), st));
    SB200_CUDA_TRY(err, cudaMemsetAsync(pair_cnt, 0, 4 * (size_t)(n_pairs + 1), st));
    if (general)
    {
        SB200_CUDA_TRY(err, cudaMallocAsync(&colj, 4 * Ts, st));
        SB200_CUDA_TRY(err, cudaMallocAsync(&w, 8 * Ts, st));
        k_emit_terms<true><<<grid_for((long long)n * 32, 256, 148 * 16), 256, 0, st>>>(
            n, csc_colptr, csc_rows, csc_vals, toff, keys, payload, colj, w, pair_cnt);
    }
    else
        k_emit_terms<false><<<grid_for((long long)n * 32, 256, 148 * 16), 256, 0, st>>>(
            n, csc_colptr, csc_rows, csc_vals, toff, keys, payload, nullptr, nullptr, pair_cnt);
    const int nb = bits_for((unsigned long long)n_pairs);
    size_t sort_bytes = 0, scan_bytes = 0;
    SB200_CUDA_TRY(err, cub::DeviceRadixSort::SortPairs(nullptr, sort_bytes, keys, keys_out, payload,
                                                        payload_out, (long long)T, 0, nb, st));
    SB200_CUDA_TRY(err, cub::DeviceScan::ExclusiveSum(nullptr, scan_bytes, pair_cnt, out->pair_ptr,
                                                      n_pairs + 1, st));
    tmp_bytes = sort_bytes > scan_bytes ? sort_bytes : scan_bytes;
    SB200_CUDA_TRY(err, cudaMallocAsync(&tmp, tmp_bytes, st));
    size_t tb = tmp_bytes;
    SB200_CUDA_TRY(err, cub::DeviceRadixSort::SortPairs(tmp, tb, keys, keys_out, payload, payload_out,
                                                        (long long)T, 0, nb, st));
    tb = tmp_bytes;
    SB200_CUDA_TRY(err, cub::DeviceScan::ExclusiveSum(tmp, tb, pair_cnt, out->pair_ptr, n_pairs + 1, st));
    if (general)
    {
        SB200_CUDA_TRY(err, cudaMallocAsync(&out->term_col, 4 * Ts, st));
        SB200_CUDA_TRY(err, cudaMallocAsync(&out->term_w, 8 * Ts, st));
        k_gather_terms<<<grid_for((long long)T, 256, 148 * 16), 256, 0, st>>>((long long)T, payload_out, colj, w,
                                                                            out->term_col, out->term_w);
        SB200_CUDA_TRY(err, cudaStreamSynchronize(st));
        cudaFreeAsync(payload_out, st);
    }
    else
    {
        SB200_CUDA_TRY(err, cudaStreamSynchronize(st));
        out->term_col = payload_out;      // the sorted payload IS the column list
        out->term_w = nullptr;
    }
    g_launch_count += 3;
    if (!general && pad_id < 65535 && T > 0)
    {   // repack: 2-byte column ids, every entry's list padded to whole 16-byte chunks with the id n
        unsigned *ccnt = pair_cnt;      // reuse
        SB200_CUDA_TRY(err, cudaMallocAsync(&out->chunk_ptr, 4 * (size_t)(n_pairs + 1), st));
        k_pair_chunk_counts<<<grid_for(n_pairs + 1, 256, 148 * 16), 256, 0, st>>>(n_pairs, out->pair_ptr, ccnt);
        size_t sb = tmp_bytes;
        SB200_CUDA_TRY(err, cub::DeviceScan::ExclusiveSum(tmp, sb, ccnt, out->chunk_ptr, n_pairs + 1, st));
        unsigned total = 0;
        SB200_CUDA_TRY(err, cudaMemcpyAsync(&total, out->chunk_ptr + n_pairs, 4, cudaMemcpyDeviceToHost, st));
        SB200_CUDA_TRY(err, cudaStreamSynchronize(st));
        SB200_CUDA_TRY(err, cudaMallocAsync(&out->term16, 16 * ((size_t)total + 1), st));
        k_pack_terms16<<<grid_for(n_pairs, 256, 148 * 16), 256, 0, st>>>(n_pairs, out->pair_ptr, out->term_col,
                                                                        out->chunk_ptr, out->term16, (unsigned short)pad_id);
        {
            const char *off = getenv("SB200_DECOLLIDE");
            if (!(off && off[0] == '0') && m > 1)
            {
                k_decollide_chunks16<<<dim3((unsigned)((m / 16 + 63) / 64 > 0 ? (m / 16 + 63) / 64 : 1), (unsigned)(m - 1)), 64, 0, st>>>(
                    m, out->chunk_ptr, out->term16, (unsigned short)pad_id);
                ++g_launch_count;
            }
        }
        SB200_CUDA_TRY(err, cudaStreamSynchronize(st));
        out->n_chunks = total;
        out->pad_id = pad_id;
        cudaFreeAsync(out->term_col, st);
        out->term_col = nullptr;
        g_launch_count += 2;
    }
    cudaFreeAsync(keys, st); cudaFreeAsync(keys_out, st); cudaFreeAsync(payload, st); cudaFreeAsync(pair_cnt, st); cudaFreeAsync(toff, st); cudaFreeAsync(tmp, st);
    if (colj) cudaFreeAsync(colj, st);
    if (w) cudaFreeAsync(w, st);
    out->m = m;
    out->n_pairs = n_pairs;
    out->n_terms = (long long)T;
    return SB200_OK;
}

// ---- pattern-only row / column lists (CompactLists) ---------------------------------------------------------
__global__ void k_list_chunk_counts(int cnt, const int *__restrict__ ptr, unsigned int *__restrict__ out)
{
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i <= cnt; i += gridDim.x * blockDim.x)
        out[i] = i < cnt ? (unsigned int)(ptr[i + 1] - ptr[i] + 7) >> 3 : 0u;
}
// a warp per list: ids narrowed to 2 bytes, the tail of the last chunk filled with `pad`
__global__ void k_pack_lists16(int cnt, const int *__restrict__ ptr, const int *__restrict__ ids,
                               const unsigned int *__restrict__ cptr, unsigned short *__restrict__ out, unsigned short pad)
{
    const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
    for (int i = blockIdx.x * wpb + (threadIdx.x >> 5); i < cnt; i += gridDim.x * wpb)
    {
        const int a = ptr[i], len = ptr[i + 1] - a;
        const size_t o = (size_t)cptr[i] << 3;
        const int padded = (int)(cptr[i + 1] - cptr[i]) << 3;
        for (int t = lane; t < padded; t += 32) out[o + t] = t < len ? (unsigned short)ids[a + t] : pad;
    }
}
__global__ void k_col_sign(int n, const int *__restrict__ colptr, const double *__restrict__ vals, double *__restrict__ sign)
{
    for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < n; j += gridDim.x * blockDim.x)
        sign[j] = colptr[j + 1] > colptr[j] ? vals[colptr[j]] : 1.0;
}

void free_compact_lists(CompactLists *p, cudaStream_t st)
{
    if (p->row_ptr) cudaFreeAsync(p->row_ptr, st);
    if (p->col_ptr) cudaFreeAsync(p->col_ptr, st);
    if (p->row16) cudaFreeAsync(p->row16, st);
    if (p->col16) cudaFreeAsync(p->col16, st);
    if (p->col_sign) cudaFreeAsync(p->col_sign, st);
    *p = CompactLists{};
}

static int pack_one(ErrorSink &err, int cnt, const int *ptr, const int *ids, unsigned short pad, unsigned int **cptr_out,
                    unsigned short **list_out, unsigned int *chunks_out, cudaStream_t st)
{
    unsigned int *counts = nullptr;
    void *tmp = nullptr;
    size_t tmp_bytes = 0;
    SB200_CUDA_TRY(err, cudaMallocAsync(&counts, 4 * (size_t)(cnt + 1), st));
    SB200_CUDA_TRY(err, cudaMallocAsync(cptr_out, 4 * (size_t)(cnt + 1), st));
    k_list_chunk_counts<<<grid_for(cnt + 1, 256, 148 * 8), 256, 0, st>>>(cnt, ptr, counts);
    SB200_CUDA_TRY(err, cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, counts, *cptr_out, cnt + 1, st));
    SB200_CUDA_TRY(err, cudaMallocAsync(&tmp, tmp_bytes, st));
    SB200_CUDA_TRY(err, cub::DeviceScan::ExclusiveSum(tmp, tmp_bytes, counts, *cptr_out, cnt + 1, st));
    unsigned int total = 0;
    SB200_CUDA_TRY(err, cudaMemcpyAsync(&total, *cptr_out + cnt, 4, cudaMemcpyDeviceToHost, st));
    SB200_CUDA_TRY(err, cudaStreamSynchronize(st));
    SB200_CUDA_TRY(err, cudaMallocAsync(list_out, 16 * ((size_t)total + 1), st));
    k_pack_lists16<<<grid_for((long long)cnt * 32, 256, 148 * 8), 256, 0, st>>>(cnt, ptr, ids, *cptr_out, *list_out, pad);
    SB200_CUDA_TRY(err, cudaGetLastError());
    cudaFreeAsync(counts, st);
    cudaFreeAsync(tmp, st);
    *chunks_out = total;
    g_launch_count += 3;
    return SB200_OK;
}

int build_compact_lists(ErrorSink &err, int m, int n, const int *csr_offs, const int *csr_inds, const int *csc_colptr,
                        const int *csc_rows, const double *csc_vals, CompactLists *out, cudaStream_t st)
{
    free_compact_lists(out, st);
    if (m >= 65535 || n >= 65535)
    {
        err.msg = "build_compact_lists: ids do not fit 2 bytes";
        return SB200_ERR_UNSUPPORTED;
    }
    int rc = pack_one(err, m, csr_offs, csr_inds, (unsigned short)n, &out->row_ptr, &out->row16, &out->row_chunks, st);
    if (rc == SB200_OK)
        rc = pack_one(err, n, csc_colptr, csc_rows, (unsigned short)m, &out->col_ptr, &out->col16, &out->col_chunks, st);
    if (rc != SB200_OK)
    {
        free_compact_lists(out, st);
        return rc;
    }
    SB200_CUDA_TRY(err, cudaMallocAsync(&out->col_sign, 8 * (size_t)(n > 0 ? n : 1), st));
    k_col_sign<<<grid_for(n, 256, 148 * 8), 256, 0, st>>>(n, csc_colptr, csc_vals, out->col_sign);
    SB200_CUDA_TRY(err, cudaGetLastError());
    ++g_launch_count;
    out->m = m;
    out->n = n;
    return SB200_OK;
}

// ---------------------------------------------------------------------------------------------
// numeric assembly: one thread per lower-triangular entry
// ---------------------------------------------------------------------------------------------
template <bool GENERAL>
__global__ void __launch_bounds__(256)
k_assemble_normal(long long n_pairs, const unsigned int *__restrict__ pair_ptr,
                  const unsigned int *__restrict__ term_col, const double *__restrict__ term_w,
                  const double *__restrict__ d, double *__restrict__ M, int ld)
{
    for (long long p = blockIdx.x * (long long)blockDim.x + threadIdx.x; p < n_pairs;
         p += (long long)gridDim.x * blockDim.x)
    {
        const unsigned int a = pair_ptr[p], e = pair_ptr[p + 1];
        double sum = 0.0;
        for (unsigned int t = a; t < e; ++t)
        {
            const double dj = __ldg(d + term_col[t]);
            sum += GENERAL ? term_w[t] * dj : dj;
        }
        long long i = (long long)((sqrt(8.0 * (double)p + 1.0) - 1.0) * 0.5);
        while ((i + 1) * (i + 2) / 2 <= p) ++i;
        while (i * (i + 1) / 2 > p) --i;
        const long long k = p - i * (i + 1) / 2;
        M[i * ld + k] = sum;
    }
}

// compact unit-product form: one thread per entry, its column ids arrive as aligned 16-byte chunks
// (three requested together: one memory latency per entry instead of one per term); d is gathered
// through the read-only path (88 KB at n = 11000: L1-resident); pad ids read d[n] = 0
__device__ __forceinline__ double chunk_gather8(uint4 v, const double *__restrict__ d)
{
    return ((__ldg(d + (v.x & 0xffffu)) + __ldg(d + (v.x >> 16))) + (__ldg(d + (v.y & 0xffffu)) + __ldg(d + (v.y >> 16)))) +
           ((__ldg(d + (v.z & 0xffffu)) + __ldg(d + (v.z >> 16))) + (__ldg(d + (v.w & 0xffffu)) + __ldg(d + (v.w >> 16))));
}
__global__ void __launch_bounds__(256)
k_assemble_normal16(long long n_pairs, const unsigned int *__restrict__ chunk_ptr, const uint4 *__restrict__ term8,
                    const double *__restrict__ d, double *__restrict__ M, int ld)
{
    for (long long p = blockIdx.x * (long long)blockDim.x + threadIdx.x; p < n_pairs;
         p += (long long)gridDim.x * blockDim.x)
    {
        const unsigned int a = chunk_ptr[p], e = chunk_ptr[p + 1];
        double s0 = 0.0, s1 = 0.0, s2 = 0.0;
        for (unsigned int c = a; c < e; c += 3)
        {
            const uint4 v0 = __ldg(term8 + c);
            uint4 v1 = make_uint4(0u, 0u, 0u, 0u), v2 = v1;
            const bool h1 = c + 1 < e, h2 = c + 2 < e;
            if (h1) v1 = __ldg(term8 + c + 1);
            if (h2) v2 = __ldg(term8 + c + 2);
            s0 += chunk_gather8(v0, d);
            if (h1) s1 += chunk_gather8(v1, d);
            if (h2) s2 += chunk_gather8(v2, d);
        }
        long long i = (long long)((sqrt(8.0 * (double)p + 1.0) - 1.0) * 0.5);
        while ((i + 1) * (i + 2) / 2 <= p) ++i;
        while (i * (i + 1) / 2 > p) --i;
        const long long k = p - i * (i + 1) / 2;
        M[i * ld + k] = (s0 + s1) + s2;
    }
}

// Same, with d staged in shared memory by persistent CTAs (one per SM).  ncu on the version above at scpnrh
// size: lg_throttle 31 / long_scoreboard 20 stall cycles per issue - every 8-byte gather of d is its own L1
// sector request (up to 32 per warp instruction), 14 M of them per launch.  From shared memory the same
// gather is ~6 wavefronts per warp instruction; d (88 KB at n = 11000) is staged once per CTA.
static constexpr int ASM_SMEM_THREADS = 1024;
__global__ void __launch_bounds__(ASM_SMEM_THREADS, 1)
k_assemble_normal16_smem(long long n_pairs, int m_rows, const unsigned int *__restrict__ chunk_ptr,
                         const uint4 *__restrict__ term8, const double *__restrict__ d, int nd, double *__restrict__ M, int ld)
{
    extern __shared__ __align__(16) double ds[];
    pdl_wait();
    for (int i = threadIdx.x; i < nd; i += ASM_SMEM_THREADS)
        ds[i] = d[i];
    __syncthreads();
    // ncu (source page) on the first form of this loop: 30 % of the stall samples sat on the first use of a
    // chunk load and a warp ran 3.2 three-chunk trips per entry although the mean list is 3.1 chunks: the
    // DIAGONAL entries carry a whole row (500 terms at scpnrh size) and one of them per row boundary kept
    // its warp - and, at the end, the whole kernel - waiting.  So: off-diagonal entries one thread each with
    // six chunk loads in flight and the next entry's bounds requested early; diagonal entries afterwards,
    // one WARP each, lanes striding the chunks.
    const long long stride = (long long)gridDim.x * ASM_SMEM_THREADS;
    long long p = blockIdx.x * (long long)ASM_SMEM_THREADS + threadIdx.x;
    unsigned int a = 0, e = 0;
    if (p < n_pairs)
    {
        a = __ldg(chunk_ptr + p);
        e = __ldg(chunk_ptr + p + 1);
    }
    while (p < n_pairs)
    {
        const long long pn = p + stride;
        unsigned int na = 0, ne = 0;
        if (pn < n_pairs)
        {
            na = __ldg(chunk_ptr + pn);
            ne = __ldg(chunk_ptr + pn + 1);
        }
        long long i = (long long)((sqrt(8.0 * (double)p + 1.0) - 1.0) * 0.5);
        while ((i + 1) * (i + 2) / 2 <= p) ++i;
        while (i * (i + 1) / 2 > p) --i;
        const long long k = p - i * (i + 1) / 2;
        if (k == i) e = a;                      // diagonal entry: left to the warp pass below
        double s0 = 0.0, s1 = 0.0, s2 = 0.0;
        for (unsigned int c = a; c < e; c += 6)
        {
            uint4 v[6];
#pragma unroll
            for (int q = 0; q < 6; ++q)
                v[q] = (c + q < e) ? __ldg(term8 + c + q) : make_uint4(0u, 0u, 0u, 0u);
            s0 += chunk_gather8_s(v[0], ds);
            if (c + 1 < e) s1 += chunk_gather8_s(v[1], ds);
            if (c + 2 < e) s2 += chunk_gather8_s(v[2], ds);
            if (c + 3 < e) s0 += chunk_gather8_s(v[3], ds);
            if (c + 4 < e) s1 += chunk_gather8_s(v[4], ds);
            if (c + 5 < e) s2 += chunk_gather8_s(v[5], ds);
        }
        if (k != i) M[i * ld + k] = (s0 + s1) + s2;
        p = pn;
        a = na;
        e = ne;
    }
    // diagonal entries: M[r][r] = sum over the columns of row r, one warp per row
    const int lane = threadIdx.x & 31, wpb = ASM_SMEM_THREADS >> 5;
    for (long long r = blockIdx.x * (long long)wpb + (threadIdx.x >> 5); r < m_rows; r += (long long)gridDim.x * wpb)
    {
        const long long pd = r * (r + 1) / 2 + r;
        const unsigned int da = __ldg(chunk_ptr + pd), de = __ldg(chunk_ptr + pd + 1);
        double sum = 0.0;
        for (unsigned int c = da + lane; c < de; c += 32)
            sum += chunk_gather8_s(__ldg(term8 + c), ds);
        sum = warp_sum(sum);
        if (lane == 0) M[r * ld + r] = sum;
    }
}

void launch_assemble_normal(const NormalPattern &P, const double *d, double *M, int ld, cudaStream_t st)
{
    const int grid = grid_for(P.n_pairs, 256, 148 * 64);
#ifndef SB200_ASM_SMEM_MIN_CTAS
#define SB200_ASM_SMEM_MIN_CTAS 32     // entries for at least this many CTAs of 1024 threads: below, staging d per CTA costs more than it saves
#endif
    if (P.term16 && P.pad_id >= 0 && (size_t)(P.pad_id + 1) * 8 <= 200 * 1024 &&
        P.n_pairs >= (long long)SB200_ASM_SMEM_MIN_CTAS * ASM_SMEM_THREADS)
    {
        static unsigned long long attr_seen = 0;
        if (first_use_on_device(attr_seen))
            cudaFuncSetAttribute(k_assemble_normal16_smem, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        const int nd = P.pad_id + 1;
        // one entry per thread and trip at least: m = 500 (scpnre / scpnrf, B&B nodes) fills 123 CTAs, m >= 550 all 148
        const int ctas = (int)std::min<long long>(148, (P.n_pairs + ASM_SMEM_THREADS - 1) / ASM_SMEM_THREADS);
        launch_pdl(k_assemble_normal16_smem, ctas, ASM_SMEM_THREADS, sizeof(double) * (size_t)nd, st,
                   P.n_pairs, P.m, P.chunk_ptr, reinterpret_cast<const uint4 *>(P.term16), d, nd, M, ld);
        ++g_launch_count;
        return;
    }
    if (P.term16)
    {
        k_assemble_normal16<<<grid, 256, 0, st>>>(P.n_pairs, P.chunk_ptr, reinterpret_cast<const uint4 *>(P.term16), d, M, ld);
        ++g_launch_count;
        return;
    }
    if (P.term_w)
        k_assemble_normal<true><<<grid, 256, 0, st>>>(P.n_pairs, P.pair_ptr, P.term_col, P.term_w, d, M, ld);
    else
        k_assemble_normal<false><<<grid, 256, 0, st>>>(P.n_pairs, P.pair_ptr, P.term_col, nullptr, d, M, ld);
    ++g_launch_count;
}

// ---------------------------------------------------------------------------------------------
// FP64 tensor-core SYRK: C(lower, 64x64 tiles) = A diag(d) A', A dense row-major (rows padded to 64,
// columns padded to 32 with zeros)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
k_syrk_dmma(int kpad, const double *__restrict__ A, int lda, const double *__restrict__ d,
            double *__restrict__ C, int ld)
{
    __shared__ __align__(16) double smem[2 * TB * KP];
    __shared__ double dsh[KC];
    double(*As)[KP] = reinterpret_cast<double(*)[KP]>(smem);
    double(*Bs)[KP] = reinterpret_cast<double(*)[KP]>(smem + TB * KP);
    const int p = blockIdx.x;
    int ti = (int)((sqrt(8.0 * p + 1.0) - 1.0) * 0.5);
    while ((ti + 1) * (ti + 2) / 2 <= p) ++ti;
    while (ti * (ti + 1) / 2 > p) --ti;
    const int tj = p - ti * (ti + 1) / 2;
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const int wm = w >> 1, wn = w & 1, g = lane >> 2, tg = lane & 3;
    const size_t r0 = (size_t)ti * TB, c0 = (size_t)tj * TB;
    double acc[4][4][2];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j)
            acc[i][j][0] = acc[i][j][1] = 0.0;
    for (int kc = 0; kc < kpad; kc += KC)
    {
        __syncthreads();
        if (tid < KC) dsh[tid] = d[kc + tid];
        __syncthreads();
        load_tile_64xKC(As, A + r0 * lda + kc, lda, tid, 128, dsh);
        load_tile_64xKC(Bs, A + c0 * lda + kc, lda, tid, 128, nullptr);
        __syncthreads();
        warp_mma_32x32(As, Bs, wm, wn, lane, 1.0, acc);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j)
            *reinterpret_cast<double2 *>(C + (r0 + wm * 32 + i * 8 + g) * ld + c0 + wn * 32 + j * 8 +
                                         tg * 2) = make_double2(acc[i][j][0], acc[i][j][1]);
}

void launch_syrk_dmma(int m, int k, const double *a, int lda, const double *d, double *c, int ld,
                      cudaStream_t st)
{
    (void)m;
    const int T = ld / TB;
    const int kpad = (k + KC - 1) / KC * KC;
    k_syrk_dmma<<<T * (T + 1) / 2, 128, 0, st>>>(kpad, a, lda, d, c, ld);
    ++g_launch_count;
}

} // namespace sb200
