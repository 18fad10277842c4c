// sb200_blocked.cu - shared-memory-staged sparse products for the large (PCG) instances.
//
// On the 50k x 1M synthetic instance the plain CSR/CSC kernels (sb200_spmv.cu) are bound by the
// gather, not by the matrix stream: every stored entry pulls a 32-byte L2 sector for 8 useful bytes of
// the dense vector (1.6 GB of sector traffic per product against 0.6 GB of matrix).  A set-covering
// matrix [A0 | -I] has only +/-1 entries, so here
//   * the matrix is stored as a PATTERN: 2 bytes per entry (15-bit column/row index local to a block of
//     the dense vector + a sign bit) instead of 12;
//   * the dense vector is cut into blocks that fit shared memory (<= 16384 doubles); a CTA stages one
//     block and serves every gather of its entries from shared memory;
//   * entries are ordered (vector block, row/column) so each CTA streams one contiguous range.
// Replaces the same reference call sites as sb200_spmv.cu (cusparseSpMV NON_TRANSPOSE / TRANSPOSE,
// /root/reference/src/sypha_solver_krylov.cu:215,309,324,428; sypha_solver.cpp:419,450) when the model
// is loaded for the PCG strategy and every coefficient is +1 or -1; other models keep the value-carrying
// kernels.  Results are deterministic (fixed summation order), sums are taken in block order.
#include "sb200_kernels.cuh"
#include "sb200_pcg.cuh"

#include <algorithm>
#include <cstdlib>
#include <cub/cub.cuh>

namespace sb200 {

static constexpr int BLK_ROW_THREADS = 1024;   // A v   : one CTA per (vector block, row chunk)
static constexpr int BLK_COL_THREADS = 512;    // A' v  : one CTA per column chunk, loops the vector blocks
static constexpr int BLK_COLS_PER_THREAD = 8;
static constexpr int BLK_COL_CHUNK = BLK_COL_THREADS * BLK_COLS_PER_THREAD;

// ---------------------------------------------------------------------------------------------
// build
// ---------------------------------------------------------------------------------------------
__global__ void k_blk_count(int majors, const int *__restrict__ mptr, const int *__restrict__ midx,
                            const double *__restrict__ vals, int nb, unsigned *__restrict__ cnt, int *__restrict__ general)
{
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < majors; i += gridDim.x * blockDim.x)
    {
        bool unit = true;
        for (int k = mptr[i]; k < mptr[i + 1]; ++k)
        {
            const double v = vals[k];
            unit = unit && (v == 1.0 || v == -1.0);
            cnt[(size_t)(midx[k] / nb) * majors + i] += 1u;       // the thread owns every counter of its major
        }
        if (!unit) atomicOr(general, 1);
    }
}
// entries -> 8-entry chunks per segment
__global__ void k_blk_to_chunks(size_t nseg, unsigned *__restrict__ cnt)
{
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < nseg; i += (size_t)gridDim.x * blockDim.x)
        cnt[i] = (cnt[i] + 7u) >> 3;
}
__global__ void k_blk_cursor(size_t nseg, const unsigned *__restrict__ ptr, unsigned *__restrict__ cursor)
{
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < nseg; i += (size_t)gridDim.x * blockDim.x)
        cursor[i] = ptr[i] << 3;
}
__global__ void k_blk_fill(int majors, const int *__restrict__ mptr, const int *__restrict__ midx,
                           const double *__restrict__ vals, int nb, unsigned *__restrict__ cursor,
                           unsigned short *__restrict__ ent)
{
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < majors; i += gridDim.x * blockDim.x)
        for (int k = mptr[i]; k < mptr[i + 1]; ++k)
        {
            const int b = midx[k] / nb;
            const unsigned pos = cursor[(size_t)b * majors + i]++;
            ent[pos] = (unsigned short)((midx[k] - b * nb) | (vals[k] < 0.0 ? 0x8000 : 0));
        }
}
// the tail of every segment's last chunk points at the zero kept behind the staged vector block
__global__ void k_blk_pad(int majors, int nblk, int nb, const unsigned *__restrict__ ptr,
                          const unsigned *__restrict__ cursor, unsigned short *__restrict__ ent)
{
    const size_t nseg = (size_t)nblk * majors;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < nseg; i += (size_t)gridDim.x * blockDim.x)
        for (unsigned pos = cursor[i]; pos < (ptr[i + 1] << 3); ++pos)
            ent[pos] = (unsigned short)nb;
}

// Both product kernels sum 16 consecutive majors side by side (a half-warp), chunk position by chunk position, every entry
// an 8-byte gather from the staged vector block: the order of the entries INSIDE a chunk is free and is chosen here so that
// the 16 lanes' entries of one slot fall into different shared-memory banks where a free slot allows it (greedy, pads
// last) - the same idea as k_decollide_chunks16 for the normal-matrix structure.
__global__ void k_blk_decollide(int majors, int nblk, const unsigned *__restrict__ ptr, unsigned short *__restrict__ ent,
                                unsigned short pad)
{
    const int groups = (majors + 15) / 16;
    const size_t total = (size_t)nblk * groups;
    for (size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x; t < total; t += (size_t)gridDim.x * blockDim.x)
    {
        const int cb = (int)(t / groups), g = (int)(t % groups);
        const int i0 = g * 16, cnt = min(16, majors - i0);
        const unsigned *p = ptr + (size_t)cb * majors + i0;
        unsigned a[16], len[16], maxlen = 0;
        for (int l = 0; l < 16; ++l)
        {
            a[l] = l < cnt ? p[l] : 0u;
            len[l] = l < cnt ? p[l + 1] - a[l] : 0u;
            maxlen = max(maxlen, len[l]);
        }
        for (unsigned j = 0; j < maxlen; ++j)
        {
            unsigned char load[8][16];                                     // ids per (slot, bank) so far
            for (int q = 0; q < 8; ++q)
                for (int b = 0; b < 16; ++b) load[q][b] = 0;
            for (int l = 0; l < cnt; ++l)
            {
                if (j >= len[l]) continue;
                uint4 *chunk = reinterpret_cast<uint4 *>(ent) + a[l] + j;
                const uint4 v = *chunk;
                const unsigned short ids[8] = {(unsigned short)(v.x & 0xffffu), (unsigned short)(v.x >> 16),
                                               (unsigned short)(v.y & 0xffffu), (unsigned short)(v.y >> 16),
                                               (unsigned short)(v.z & 0xffffu), (unsigned short)(v.z >> 16),
                                               (unsigned short)(v.w & 0xffffu), (unsigned short)(v.w >> 16)};
                unsigned short out[8];
                unsigned used = 0;
                for (int q8 = 0; q8 < 8; ++q8)
                {
                    const unsigned short id = ids[q8];
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
                *chunk = make_uint4((unsigned)out[0] | ((unsigned)out[1] << 16), (unsigned)out[2] | ((unsigned)out[3] << 16),
                                    (unsigned)out[4] | ((unsigned)out[5] << 16), (unsigned)out[6] | ((unsigned)out[7] << 16));
            }
        }
    }
}

void free_blocked(BlockedPattern *p, cudaStream_t st)
{
    if (p->ptr) cudaFreeAsync(p->ptr, st);
    if (p->ent) cudaFreeAsync(p->ent, st);
    if (p->partial) cudaFreeAsync(p->partial, st);
    *p = BlockedPattern{};
}

// returns SB200_ERR_UNSUPPORTED (and leaves *out empty) when some coefficient is not +/-1
int build_blocked(ErrorSink &err, int majors, int minors, long long nnz, const int *mptr, const int *midx,
                  const double *vals, int nb, bool with_partials, BlockedPattern *out, cudaStream_t st)
{
    free_blocked(out, st);
    const int nblk = (minors + nb - 1) / nb;
    const size_t nseg = (size_t)nblk * majors;
    unsigned *cnt = nullptr, *cursor = nullptr;
    int *general = nullptr;
    void *tmp = nullptr;
    size_t tmp_bytes = 0;
    SB200_CUDA_TRY(err, cudaMallocAsync(&cnt, sizeof(unsigned) * (nseg + 1), st));
    SB200_CUDA_TRY(err, cudaMallocAsync(&cursor, sizeof(unsigned) * (nseg + 1), st));
    SB200_CUDA_TRY(err, cudaMallocAsync(&general, sizeof(int), st));
    SB200_CUDA_TRY(err, cudaMemsetAsync(cnt, 0, sizeof(unsigned) * (nseg + 1), st));
    SB200_CUDA_TRY(err, cudaMemsetAsync(general, 0, sizeof(int), st));
    k_blk_count<<<grid_for(majors, 128, 148 * 32), 128, 0, st>>>(majors, mptr, midx, vals, nb, cnt, general);
    int h_general = 0;
    SB200_CUDA_TRY(err, cudaMemcpyAsync(&h_general, general, sizeof(int), cudaMemcpyDeviceToHost, st));
    SB200_CUDA_TRY(err, cudaStreamSynchronize(st));
    if (h_general)
    {
        cudaFreeAsync(cnt, st); cudaFreeAsync(cursor, st); cudaFreeAsync(general, st);
        return SB200_ERR_UNSUPPORTED;
    }
    SB200_CUDA_TRY(err, cudaMallocAsync(&out->ptr, sizeof(unsigned) * (nseg + 1), st));
    k_blk_to_chunks<<<grid_for((long long)nseg, 256, 148 * 16), 256, 0, st>>>(nseg, cnt);
    SB200_CUDA_TRY(err, cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, cnt, out->ptr, nseg + 1, st));
    SB200_CUDA_TRY(err, cudaMallocAsync(&tmp, tmp_bytes, st));
    SB200_CUDA_TRY(err, cub::DeviceScan::ExclusiveSum(tmp, tmp_bytes, cnt, out->ptr, nseg + 1, st));
    unsigned total_chunks = 0;
    SB200_CUDA_TRY(err, cudaMemcpyAsync(&total_chunks, out->ptr + nseg, sizeof(unsigned), cudaMemcpyDeviceToHost, st));
    SB200_CUDA_TRY(err, cudaStreamSynchronize(st));
    out->n_chunks = total_chunks;
    SB200_CUDA_TRY(err, cudaMallocAsync(&out->ent, sizeof(uint4) * ((size_t)total_chunks + 1), st));
    k_blk_cursor<<<grid_for((long long)nseg, 256, 148 * 16), 256, 0, st>>>(nseg, out->ptr, cursor);
    k_blk_fill<<<grid_for(majors, 128, 148 * 32), 128, 0, st>>>(majors, mptr, midx, vals, nb, cursor, out->ent);
    k_blk_pad<<<grid_for((long long)nseg, 256, 148 * 16), 256, 0, st>>>(majors, nblk, nb, out->ptr, cursor, out->ent);
    g_launch_count += 5;
    {
        const char *off = getenv("SB200_DECOLLIDE");
        if (!(off && off[0] == '0'))
        {
            k_blk_decollide<<<grid_for((long long)nblk * ((majors + 15) / 16), 128, 148 * 32), 128, 0, st>>>(
                majors, nblk, out->ptr, out->ent, (unsigned short)nb);
            ++g_launch_count;
        }
    }
    if (with_partials) SB200_CUDA_TRY(err, cudaMallocAsync(&out->partial, sizeof(double) * nseg, st));
    SB200_CUDA_TRY(err, cudaStreamSynchronize(st));
    cudaFreeAsync(cnt, st); cudaFreeAsync(cursor, st); cudaFreeAsync(general, st); cudaFreeAsync(tmp, st);
    out->majors = majors;
    out->minors = minors;
    out->nb = nb;
    out->nblk = nblk;
    out->nnz = nnz;
    return SB200_OK;
}

// ---------------------------------------------------------------------------------------------
// A v (rows are the majors): partial[cb][row] = sum over the entries of row in vector block cb
// ---------------------------------------------------------------------------------------------
// Sum of one (major, vector block) segment, ONE thread per segment.  A segment is a run of whole
// 16-byte chunks (8 two-byte entries; the build pads the last chunk with entries that point at a zero
// kept behind the staged block), so there is no per-entry bounds logic and the loads depend only on the
// segment bounds: up to three chunks are requested together, one memory latency per segment.
// (History, ncu on the 50k x 1M instance: a 2-byte-load loop stalled the warp on some lane's sector
// miss in nearly every iteration - 857 cycles per entry; predicated unaligned chunks were issue-bound at
// 27 thread-instructions per entry.)
template <bool USE_SIGN>
__device__ __forceinline__ void blk_slot(unsigned half, const double *xs, double &acc)
{
    const double x = xs[half & 0x7fffu];
    if (USE_SIGN)
        acc += __hiloint2double(__double2hiint(x) ^ (int)((half & 0x8000u) << 16), __double2loint(x));
    else
        acc += x;
}
template <bool USE_SIGN>
__device__ __forceinline__ void blk_chunk(uint4 v, const double *xs, double &acc0, double &acc1)
{
    blk_slot<USE_SIGN>(v.x & 0xffffu, xs, acc0);
    blk_slot<USE_SIGN>(v.x >> 16, xs, acc1);
    blk_slot<USE_SIGN>(v.y & 0xffffu, xs, acc0);
    blk_slot<USE_SIGN>(v.y >> 16, xs, acc1);
    blk_slot<USE_SIGN>(v.z & 0xffffu, xs, acc0);
    blk_slot<USE_SIGN>(v.z >> 16, xs, acc1);
    blk_slot<USE_SIGN>(v.w & 0xffffu, xs, acc0);
    blk_slot<USE_SIGN>(v.w >> 16, xs, acc1);
}
template <bool USE_SIGN>
__device__ __forceinline__ double blk_segment_sum(const uint4 *__restrict__ ent8, unsigned a, unsigned e, const double *xs)
{
    double acc0 = 0.0, acc1 = 0.0;
    for (unsigned ch = a; ch < e; ch += 3)
    {
        const uint4 v0 = __ldg(ent8 + ch);
        uint4 v1 = make_uint4(0u, 0u, 0u, 0u), v2 = v1;
        if (ch + 1 < e) v1 = __ldg(ent8 + ch + 1);
        if (ch + 2 < e) v2 = __ldg(ent8 + ch + 2);
        blk_chunk<USE_SIGN>(v0, xs, acc0, acc1);
        if (ch + 1 < e) blk_chunk<USE_SIGN>(v1, xs, acc0, acc1);
        if (ch + 2 < e) blk_chunk<USE_SIGN>(v2, xs, acc0, acc1);
    }
    return acc0 + acc1;
}

template <bool ABS>     // ABS: sum |a| v  (= sum a^2 v for +/-1 entries: the Jacobi diagonal)
__global__ void __launch_bounds__(BLK_ROW_THREADS, 1)
k_blk_rows(BlockedPattern B, const double *__restrict__ v, int rows_per_chunk, const int *__restrict__ skip)
{
    extern __shared__ __align__(16) double xs[];
    if (skip && *skip) return;
    const int cb = blockIdx.x, tid = threadIdx.x;
    const int base = cb * B.nb, width = min(B.nb, B.minors - base);
    for (int i = tid; i < width; i += BLK_ROW_THREADS)
        xs[i] = v[base + i];
    if (tid == 0) xs[B.nb] = 0.0;
    __syncthreads();
    const int r_lo = blockIdx.y * rows_per_chunk, r_hi = min(B.majors, r_lo + rows_per_chunk);
    const unsigned *__restrict__ ptr = B.ptr + (size_t)cb * B.majors;
    const uint4 *ent8 = reinterpret_cast<const uint4 *>(B.ent);
    double *out = B.partial + (size_t)cb * B.majors;
    int row = r_lo + tid;
    unsigned a = 0, e = 0;
    if (row < r_hi)
    {
        a = __ldg(ptr + row);
        e = __ldg(ptr + row + 1);
    }
    while (row < r_hi)
    {   // the next segment's bounds are requested before this segment's entries are consumed
        const int nrow = row + BLK_ROW_THREADS;
        unsigned na = 0, ne = 0;
        if (nrow < r_hi)
        {
            na = __ldg(ptr + nrow);
            ne = __ldg(ptr + nrow + 1);
        }
        out[row] = blk_segment_sum<!ABS>(ent8, a, e, xs);
        row = nrow;
        a = na;
        e = ne;
    }
}

// out[row] = epilogue(sum_cb partial[cb][row]);  EPI 0: alpha*s + beta*z, 1: s, 2: CG (Ap = s, p.Ap)
template <int EPI>
__global__ void __launch_bounds__(256)
k_blk_rows_reduce(BlockedPattern B, const double *__restrict__ z, double *__restrict__ out, double alpha, double beta,
                  PcgVecs C, Scalars *sc)
{
    __shared__ double sh[32];
    if (EPI == 2)
        if (sc->cg_done) return;
    double dot = 0.0;
    // 8 lanes per row: lane l takes the vector blocks l, l + 8, ... (65 strided reads per row at 50k x 1M; one thread per
    // row summed them as four chains of 16 dependent latencies: 33.6 us for 26 MB that sit in L2), then a fixed-order
    // shuffle reduction inside the group
    const int gl = threadIdx.x & 7;
    const int groups = (gridDim.x * blockDim.x) >> 3;
    const int rows_round = (B.majors + groups - 1) / groups * groups;          // whole warps stay in the loop (shuffles)
    for (int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 3; row < rows_round; row += groups)
    {
        double s0 = 0.0, s1 = 0.0;
        if (row < B.majors)
        {
            int cb = gl;
            for (; cb + 8 < B.nblk; cb += 16)
            {
                s0 += B.partial[(size_t)cb * B.majors + row];
                s1 += B.partial[(size_t)(cb + 8) * B.majors + row];
            }
            if (cb < B.nblk) s0 += B.partial[(size_t)cb * B.majors + row];
        }
        double s = s0 + s1;
        s += __shfl_xor_sync(0xffffffffu, s, 1);
        s += __shfl_xor_sync(0xffffffffu, s, 2);
        s += __shfl_xor_sync(0xffffffffu, s, 4);
        if (gl != 0 || row >= B.majors) continue;
        if (EPI == 0)
            out[row] = (beta == 0.0) ? alpha * s : alpha * s + beta * z[row];
        else if (EPI == 1)
            out[row] = s;
        else
        {
            C.Ap[row] = s;
            dot += C.p[row] * s;
        }
    }
    if (EPI == 2)
    {
        dot = block_sum(dot, sh);
        if (threadIdx.x == 0) C.partial[blockIdx.x] = dot;
        if (last_block_arrives(&sc->ticket[5], gridDim.x))
        {
            const double pap = reduce_partials(C.partial, gridDim.x, sh);
            if (threadIdx.x == 0)
            {
                sc->cg_pap = pap;
                if (!(pap > 0.0) || !isfinite(pap))     // krylov.cu:335-339
                {
                    sc->cg_fail = 1;
                    sc->cg_done = 1;
                }
            }
        }
    }
}

static void launch_blk_rows(const BlockedPattern &B, bool abs_mode, const double *v, const int *skip, cudaStream_t st)
{
    static unsigned long long attr_seen = 0;
    if (first_use_on_device(attr_seen))
    {
        cudaFuncSetAttribute(k_blk_rows<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 16386 * 8);
        cudaFuncSetAttribute(k_blk_rows<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 16386 * 8);
    }
    // row chunks: (vector blocks x chunks) CTAs at one CTA per SM (the staged block takes 128 KB).  About four waves of
    // the 148 SMs, so that the last, partial wave costs little (65 blocks x 3 chunks = 195 CTAs ran as 148 + 47: the GPU
    // a third full for half of the kernel), but never fewer than two segments per thread
    int chunks = (148 * 4) / B.nblk;
    const char *cw = getenv("SB200_BLK_ROW_WAVES");
    if (cw && atoi(cw) > 0) chunks = (148 * atoi(cw)) / B.nblk;
    chunks = std::min(chunks, B.majors / (2 * BLK_ROW_THREADS));
    if (chunks < 1) chunks = 1;
    int rows_per_chunk = (B.majors + chunks - 1) / chunks;
    rows_per_chunk = (rows_per_chunk + 31) / 32 * 32;
    chunks = (B.majors + rows_per_chunk - 1) / rows_per_chunk;
    const dim3 grid(B.nblk, chunks);
    const size_t smem = sizeof(double) * ((size_t)B.nb + 2);
    if (abs_mode)
        k_blk_rows<true><<<grid, BLK_ROW_THREADS, smem, st>>>(B, v, rows_per_chunk, skip);
    else
        k_blk_rows<false><<<grid, BLK_ROW_THREADS, smem, st>>>(B, v, rows_per_chunk, skip);
    ++g_launch_count;
}

void launch_blk_spmv_rows(const BlockedPattern &B, const double *x, const double *z, double *out, double alpha,
                          double beta, cudaStream_t st)
{
    launch_blk_rows(B, false, x, nullptr, st);
    k_blk_rows_reduce<0><<<grid_for((long long)B.majors * 8, 256, 148 * 8), 256, 0, st>>>(B, z, out, alpha, beta, PcgVecs{}, nullptr);
    ++g_launch_count;
}
void launch_blk_jacobi_diag(const BlockedPattern &B, const double *d, double *diag, cudaStream_t st)
{
    launch_blk_rows(B, true, d, nullptr, st);
    k_blk_rows_reduce<1><<<grid_for((long long)B.majors * 8, 256, 148 * 8), 256, 0, st>>>(B, nullptr, diag, 1.0, 0.0, PcgVecs{}, nullptr);
    ++g_launch_count;
}
void launch_blk_cg_matvec(const BlockedPattern &B, const PcgVecs &C, Scalars *sc, cudaStream_t st)
{
    launch_blk_rows(B, false, C.q, &sc->cg_done, st);
    k_blk_rows_reduce<2><<<grid_for((long long)B.majors * 8, 256, 148 * 8), 256, 0, st>>>(B, nullptr, nullptr, 1.0, 0.0, C, sc);
    ++g_launch_count;
}

// ---------------------------------------------------------------------------------------------
// A' v (columns are the majors): a CTA owns a chunk of columns, keeps their sums in registers and
// loops over the blocks of v, staging each in shared memory; fused epilogues as in k_spmv_csc.
// ---------------------------------------------------------------------------------------------
enum { BLK_CG = 16 };   // epilogue: q = dscale .* w (PCG), besides the CSC_* modes

template <int MODE>
__global__ void __launch_bounds__(BLK_COL_THREADS, 2)
k_blk_cols(BlockedPattern B, const double *__restrict__ v, const double *__restrict__ z, double *__restrict__ out,
           double alpha, double beta, IpmVecs V, const double *__restrict__ dscale, const Scalars *sc)
{
    extern __shared__ __align__(16) double xs[];
    __shared__ double sh[32];
    if (MODE == CSC_RECOVER)
        if (V.sc->done) return;
    if (MODE == BLK_CG)
        if (sc->cg_done) return;
    const int tid = threadIdx.x;
    double m0 = DBL_MAX, m1 = DBL_MAX;
    const uint4 *ent8 = reinterpret_cast<const uint4 *>(B.ent);
    const int nchunks = (B.majors + BLK_COL_CHUNK - 1) / BLK_COL_CHUNK;
    for (int chunk = blockIdx.x; chunk < nchunks; chunk += gridDim.x)
    {
        const int c_lo = chunk * BLK_COL_CHUNK;
        double acc[BLK_COLS_PER_THREAD];
#pragma unroll
        for (int u = 0; u < BLK_COLS_PER_THREAD; ++u)
            acc[u] = 0.0;
        for (int rb = 0; rb < B.nblk; ++rb)
        {
            const int base = rb * B.nb, width = min(B.nb, B.minors - base);
            const unsigned *__restrict__ ptr = B.ptr + (size_t)rb * B.majors;
            // segment bounds first: their latency overlaps the staging of the vector block
            unsigned sa[BLK_COLS_PER_THREAD], se[BLK_COLS_PER_THREAD];
#pragma unroll
            for (int u = 0; u < BLK_COLS_PER_THREAD; ++u)
            {
                const int col = c_lo + tid + u * BLK_COL_THREADS;
                sa[u] = se[u] = 0u;
                if (col < B.majors)
                {
                    sa[u] = __ldg(ptr + col);
                    se[u] = __ldg(ptr + col + 1);
                }
            }
            __syncthreads();
            for (int i = tid; i < width; i += BLK_COL_THREADS)
                xs[i] = v[base + i];
            if (tid == 0) xs[B.nb] = 0.0;
            __syncthreads();
#pragma unroll
            for (int u = 0; u < BLK_COLS_PER_THREAD; ++u)
                acc[u] += blk_segment_sum<true>(ent8, sa[u], se[u], xs);
        }
#pragma unroll
        for (int u = 0; u < BLK_COLS_PER_THREAD; ++u)
        {
            const int col = c_lo + tid + u * BLK_COL_THREADS;
            if (col >= B.majors) continue;
            const double w = acc[u];
            if (MODE == CSC_PLAIN)
                out[col] = (beta == 0.0) ? alpha * w : alpha * w + beta * z[col];
            else if (MODE == BLK_CG)
                out[col] = dscale ? dscale[col] * w : w;
            else if (MODE == CSC_RECOVER)
            {   // krylov.cu:74-82 + utils.cu:68-79
                const double ds = V.resC[col] - w;
                const double xj = V.x[col], sj = V.s[col];
                const double dx = (V.resXS[col] - xj * ds) / sj;
                V.ds[col] = ds;
                V.dx[col] = dx;
                if (dx < 0.0) m0 = fmin(m0, -xj / dx);
                if (ds < 0.0) m1 = fmin(m1, -sj / ds);
            }
            else if (MODE == CSC_START_X)
            {
                V.x[col] = w;
                m0 = fmin(m0, w);
            }
            else if (MODE == CSC_START_S)
            {
                const double sj = V.c[col] - w;
                V.s[col] = sj;
                m1 = fmin(m1, sj);
            }
            else if (MODE == CSC_RESC)
                V.resC[col] = V.c[col] - V.s[col] - w;
        }
    }
    if (MODE == CSC_RECOVER || MODE == CSC_START_X || MODE == CSC_START_S)
    {
        m0 = block_min(m0, sh);
        m1 = block_min(m1, sh);
        if (threadIdx.x == 0)
        {
            if (MODE == CSC_RECOVER)
            {
                atomicMin(&V.sc->amax_p, ord_encode(m0));
                atomicMin(&V.sc->amax_d, ord_encode(m1));
            }
            else if (MODE == CSC_START_X)
                atomicMin(&V.sc->min_x, ord_encode(m0));
            else
                atomicMin(&V.sc->min_s, ord_encode(m1));
        }
    }
}

template <int MODE>
static void launch_blk_cols_mode(const BlockedPattern &B, const double *v, const double *z, double *out, double alpha,
                                 double beta, const IpmVecs &V, const double *dscale, const Scalars *sc, cudaStream_t st)
{
    static unsigned long long attr_seen = 0;
    if (first_use_on_device(attr_seen))
        cudaFuncSetAttribute(k_blk_cols<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 13314 * 8);
    const int nchunks = (B.majors + BLK_COL_CHUNK - 1) / BLK_COL_CHUNK;
    const int grid = nchunks < 148 * 2 ? nchunks : 148 * 2;
    k_blk_cols<MODE><<<grid, BLK_COL_THREADS, sizeof(double) * ((size_t)B.nb + 2), st>>>(B, v, z, out, alpha, beta, V, dscale, sc);
    ++g_launch_count;
}

void launch_blk_spmv_cols(const BlockedPattern &B, int mode, const double *v, const double *z, double *out,
                          double alpha, double beta, const IpmVecs *Vp, cudaStream_t st)
{
    IpmVecs V{};
    if (Vp) V = *Vp;
    switch (mode)
    {
    case CSC_PLAIN: launch_blk_cols_mode<CSC_PLAIN>(B, v, z, out, alpha, beta, V, nullptr, nullptr, st); break;
    case CSC_RECOVER: launch_blk_cols_mode<CSC_RECOVER>(B, v, z, out, alpha, beta, V, nullptr, nullptr, st); break;
    case CSC_START_X: launch_blk_cols_mode<CSC_START_X>(B, v, z, out, alpha, beta, V, nullptr, nullptr, st); break;
    case CSC_START_S: launch_blk_cols_mode<CSC_START_S>(B, v, z, out, alpha, beta, V, nullptr, nullptr, st); break;
    case CSC_RESC: launch_blk_cols_mode<CSC_RESC>(B, v, z, out, alpha, beta, V, nullptr, nullptr, st); break;
    }
}
void launch_blk_cg_cols(const BlockedPattern &B, const double *p, double *q, const double *dscale, const Scalars *sc,
                        cudaStream_t st)
{
    launch_blk_cols_mode<BLK_CG>(B, p, nullptr, q, 1.0, 0.0, IpmVecs{}, dscale, sc, st);
}

// vector-block sizes: as large as shared memory allows, equal-sized blocks
int blocked_nb_for_rows(int n)      // A v: 1 CTA / SM, up to 16384 doubles
{
    const int nblk = (n + 16383) / 16384;
    return ((n + nblk - 1) / nblk + 63) / 64 * 64;
}
int blocked_nb_for_cols(int m)      // A' v: 2 CTAs / SM, up to 13312 doubles
{
    const int nblk = (m + 13311) / 13312;
    return ((m + nblk - 1) / nblk + 63) / 64 * 64;
}

} // namespace sb200
