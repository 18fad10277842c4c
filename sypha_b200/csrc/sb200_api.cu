// sb200_api.cu - the C ABI (include/sypha_b200.h): persistent workspace, model upload and the
// Mehrotra predictor-corrector loop in normal-equations form.
//
// Host orchestration of /root/reference/src/sypha_solver.cpp:42-886 (solver_sparse_mehrotra_run):
//   starting point   init.cpp:543-652   -> GPU, same kernels as the loop with D = I
//   initial residuals solver.cpp:375-459 -> 2 fused SpMV launches + 1 reduction
//   main loop        solver.cpp:496-772 -> per iteration: assemble, factor, 2 x (rhs SpMV, solve,
//                                          recover+ratio test), 3 fused vector kernels; the only
//                                          host traffic is one read of the scalar block per poll.
// The loop body is replayed from a CUDA graph; every kernel starts with `if (done) return`, so
// iterations may be enqueued ahead of the host's convergence poll without changing the result.
#include "sb200_kernels.cuh"
#include "sb200_chol.cuh"
#include "sb200_pcg.cuh"
#include "sb200_heur.cuh"
#include "sb200_cta.cuh"

#include <algorithm>
#include <chrono>
#include <map>
#include <thread>
#include <cmath>
#include <cstdlib>
#include <cstring>

using namespace sb200;

struct sb200_ws
{
    int device = 0;
    cudaStream_t stream = nullptr;
    ErrorSink err;

    // model
    bool loaded = false;
    int m = 0, n = 0, n_orig = 0, mpad = 0;
    long long nnz = 0;
    int strategy = SB200_STRATEGY_AUTO;
    int *csr_offs = nullptr, *csr_inds = nullptr;
    double *csr_vals = nullptr;
    int *csc_colptr = nullptr, *csc_rows = nullptr;
    double *csc_vals = nullptr;
    double *c = nullptr, *b = nullptr;
    int csc_lanes = 1;
    int m_cap = 0, n_cap = 0;
    long long nnz_cap = 0;

    NormalPattern pat;
    CompactLists lists;                     // pattern-only rows / columns for the one-block solver's products
    BlockedPattern blk_rows, blk_cols;    // PCG strategy on a +/-1 matrix: shared-memory-staged products
    double *denseA = nullptr;     // SYRK strategy: dense row-major copy of A, mpad x kpad
    int kpad = 0;
    double *M = nullptr;          // mpad x mpad
    long long M_cap = 0;
    CholWork chol;

    // iterates and scratch (one slab)
    double *slab = nullptr;
    size_t slab_bytes = 0;
    IpmVecs V{};
    double *ones_n = nullptr;
    double *cg_diag = nullptr, *cg_x = nullptr, *cg_r = nullptr, *cg_z = nullptr, *cg_p = nullptr,
           *cg_Ap = nullptr, *cg_q = nullptr;
    Scalars *sc = nullptr;
    DevParams *dparams = nullptr;
    Scalars *sc_host = nullptr;   // pinned
    DevParams hparams{};

    // fingerprint of the last loaded model (dims, CSR pattern and values): an identical re-load keeps the CSC copy, the
    // symbolic structure of M, the dense copy and the captured iteration graph (1.2 ms of radix sorts + a graph
    // instantiation per load otherwise - every e2e step, every re-solve of the same pattern)
    unsigned long long fp[2] = {0, 0};
    unsigned long long *fp_dev = nullptr;
    bool fp_valid = false;
    int fp_strategy_hint = -1;

    // B&B node = base model + node_k appended branch rows (build_branch_model, bnb.cpp:453-468), folded on
    // the device: CSR rows appended in place, CSC rebuilt from a kept copy of the base CSC, the symbolic
    // structure of M reused for the base rows and the extra rows of M written by their own kernel
    int base_m = 0, base_n = 0;
    long long base_nnz = 0;
    int node_k = 0;
    int *base_colptr = nullptr, *base_rows = nullptr;
    double *base_cvals = nullptr;
    bool base_csc_valid = false;
    int *d_var = nullptr;
    double *d_coef = nullptr;
    unsigned char *h_delta = nullptr;      // pinned staging: var | coef | rhs
    int delta_cap = 0;
    std::map<int, std::pair<cudaGraphExec_t, long long>> node_graphs;   // iteration graph per depth
    // per-node branching / incumbent kernel (sb200_heur.cu)
    int *heur_list = nullptr, *heur_sorted = nullptr;
    unsigned char *heur_cover = nullptr, *heur_nif = nullptr;
    double *heur_score = nullptr;
    int heur_rules = SB200_HEUR_REFERENCE, heur_branch_rule = SB200_BRANCH_MOST_FRACTIONAL;
    double heur_tol = 1e-6;                 // kBnbIntegralityTol
    sb200_heur_result *heur_out = nullptr, *heur_out_host = nullptr;   // device (unused since the kernel writes the pinned record), pinned
    int *heur_flag_host = nullptr;          // pinned: the node kernel sets it to heur_seq when its record is complete
    int heur_seq = 0;
    int heur_cap = 0;

    // graph of one IPM iteration (direct strategies)
    cudaGraphExec_t iter_graph = nullptr;
    long long iter_graph_kernels = 0;
    cudaGraphExec_t cg_graph = nullptr;     // a chunk of CG iterations
    long long cg_graph_kernels = 0;
    int cg_graph_chunk = 0;
    const double *cg_graph_dscale = nullptr;

    // throughput form: the whole LP by one CTA in one launch (sb200_cta.cu)
    int solver_form = SB200_FORM_LATENCY;
    CtaLp *cta_dev = nullptr, *cta_host = nullptr;     // device, pinned
    bool cta_launched = false;
    CtaLp *batch_dev = nullptr, *batch_host = nullptr;      // a window of LPs in one launch (owned by the window's first slot)
    HeurArgs *hbatch_dev = nullptr, *hbatch_host = nullptr;
    int batch_cap = 0;
    unsigned char *nd_host = nullptr, *nd_dev = nullptr;   // a window's node deltas: descriptors | var | coef | rhs (lead only)
    size_t nd_cap = 0;
    bool window_pending = false;            // sb200_window_begin without its sb200_window_finish yet
    double window_ms = 0.0;                 // device time (CUDA events on the launching stream) of the last one-launch window
    int window_lps = 0;
    const double *warm_ptr = nullptr;                  // parent's x | y | s for the next solve (one use)
    int warm_n = 0, warm_m = 0;
    double warm_floor = 0.1;

    // async solve state
    sb200_params params{};
    bool active = false;
    int enqueued = 0;
    cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
    long long launches_at_begin = 0;
    long long graph_kernel_launches = 0;
    std::vector<double> trace_host;
    int trace_rows = 0;
    bool trace_stale = false;
    int base_n_or_n() const { return node_k ? base_n : n; }
    int base_m_or_m() const { return node_k ? base_m : m; }
};

namespace {

inline int round_up(int v, int q) { return (v + q - 1) / q * q; }

#define WS_TRY(call) SB200_CUDA_TRY(ws->err, call)

int fail(sb200_ws *ws, int code, const char *msg)
{
    ws->err.msg = msg;
    return code;
}

void drop_graphs(sb200_ws *ws)
{
    for (auto &kv : ws->node_graphs)
        if (kv.second.first && kv.second.first != ws->iter_graph) cudaGraphExecDestroy(kv.second.first);
    ws->node_graphs.clear();
    if (ws->iter_graph) cudaGraphExecDestroy(ws->iter_graph);
    if (ws->cg_graph) cudaGraphExecDestroy(ws->cg_graph);
    ws->iter_graph = nullptr;
    ws->cg_graph = nullptr;
}

template <typename T>
int grow(sb200_ws *ws, T **p, size_t count)
{
    if (*p) cudaFree(*p);
    *p = nullptr;
    WS_TRY(cudaMalloc(p, sizeof(T) * (count ? count : 1)));
    return SB200_OK;
}

int ensure_capacity(sb200_ws *ws, int m, int n, long long nnz)
{
    if (nnz > ws->nnz_cap)
    {
        int rc;
        if ((rc = grow(ws, &ws->csr_inds, (size_t)nnz))) return rc;
        if ((rc = grow(ws, &ws->csr_vals, (size_t)nnz))) return rc;
        if ((rc = grow(ws, &ws->csc_rows, (size_t)nnz))) return rc;
        if ((rc = grow(ws, &ws->csc_vals, (size_t)nnz))) return rc;
        ws->nnz_cap = nnz;
    }
    if (m > ws->m_cap || n > ws->n_cap)
    {
        const int mc = std::max(m, ws->m_cap), nc = std::max(n, ws->n_cap);
        const int mp = round_up(mc, SB200_TILE);
        int rc;
        if ((rc = grow(ws, &ws->csr_offs, (size_t)mc + 1))) return rc;
        if ((rc = grow(ws, &ws->csc_colptr, (size_t)nc + 1))) return rc;
        if ((rc = grow(ws, &ws->c, (size_t)nc))) return rc;
        if ((rc = grow(ws, &ws->b, (size_t)mc))) return rc;
        // slab: 10 n-vectors, 9 m-vectors (padded), partials, trace
        const size_t nv = (size_t)round_up(nc + 1, 32), mv = (size_t)mp;     // +1: d[n] = 0 pad slot of the compact assembly
        const size_t doubles = 11 * nv + 10 * mv + 4 * (size_t)SB200_MAX_PARTIAL_BLOCKS +
                               (size_t)SB200_TRACE_ROWS * SB200_TRACE_COLS;
        if (ws->slab) cudaFree(ws->slab);
        ws->slab = nullptr;
        WS_TRY(cudaMalloc(&ws->slab, doubles * sizeof(double)));
        ws->slab_bytes = doubles * sizeof(double);
        ws->m_cap = mc;
        ws->n_cap = nc;
    }
    return SB200_OK;
}

void carve(sb200_ws *ws)
{
    const size_t nv = (size_t)round_up(ws->n_cap + 1, 32), mv = (size_t)round_up(ws->m_cap, SB200_TILE);
    double *p = ws->slab;
    auto take = [&](size_t k) { double *r = p; p += k; return r; };
    IpmVecs &V = ws->V;
    V.m = ws->m; V.n = ws->n; V.n_orig = ws->n_orig; V.mpad = ws->mpad;
    V.c = ws->c; V.b = ws->b;
    V.x = take(nv); V.s = take(nv); V.dx = take(nv); V.ds = take(nv);
    V.resC = take(nv); V.resXS = take(nv); V.d = take(nv); V.t = take(nv);
    ws->ones_n = take(nv); ws->cg_q = take(nv);
    take(nv);
    V.y = take(mv); V.resB = take(mv); V.rhs = take(mv);
    ws->cg_diag = take(mv); ws->cg_x = take(mv); ws->cg_r = take(mv); ws->cg_z = take(mv);
    ws->cg_p = take(mv); ws->cg_Ap = take(mv);
    take(mv);
    V.partial = take(4 * (size_t)SB200_MAX_PARTIAL_BLOCKS);
    V.trace = take((size_t)SB200_TRACE_ROWS * SB200_TRACE_COLS);
    V.sc = ws->sc;
    V.dy = (ws->strategy == SB200_STRATEGY_PCG) ? ws->cg_x : V.rhs;
}

CsrView csr_of(const sb200_ws *ws)
{
    return CsrView{ws->m, ws->csr_offs, ws->csr_inds, ws->csr_vals, ws->blk_rows.ptr ? &ws->blk_rows : nullptr};
}
CscView csc_of(const sb200_ws *ws)
{
    return CscView{ws->n, ws->csc_colptr, ws->csc_rows, ws->csc_vals, ws->csc_lanes,
                   ws->blk_cols.ptr ? &ws->blk_cols : nullptr};
}

__global__ void k_densify(int m, const int *__restrict__ offs, const int *__restrict__ inds,
                          const double *__restrict__ vals, double *__restrict__ A, int lda)
{
    const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
    for (int row = blockIdx.x * wpb + (threadIdx.x >> 5); row < m; row += gridDim.x * wpb)
        for (int k = offs[row] + lane; k < offs[row + 1]; k += 32)
            atomicAdd(&A[(size_t)row * lda + inds[k]], vals[k]);   // duplicates sum, like CSR semantics
}

// a (row, column) pair stored twice: the products sum duplicates, the symbolic structure of M counts a column's rows once
// each, so the two would disagree - such a model is rejected at load.  CSC rows are ascending within a column.
__global__ void k_find_duplicates(int n, const int *__restrict__ colptr, const int *__restrict__ rows, int *__restrict__ flag)
{
    for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < n; j += gridDim.x * blockDim.x)
        for (int k = colptr[j] + 1; k < colptr[j + 1]; ++k)
            if (rows[k] == rows[k - 1]) *flag = 1;
}

// ---- model fingerprint ---------------------------------------------------------------------------------
__device__ __forceinline__ unsigned long long mix64(unsigned long long v)
{
    v ^= v >> 33; v *= 0xff51afd7ed558ccdull; v ^= v >> 33; v *= 0xc4ceb9fe1a85ec53ull; v ^= v >> 33;
    return v;
}
// two independent position-sensitive 64-bit sums over offs | inds | vals (order of accumulation irrelevant: sums)
__global__ void k_fingerprint(int m, long long nnz, const int *__restrict__ offs, const int *__restrict__ inds,
                              const double *__restrict__ vals, unsigned long long *__restrict__ out)
{
    unsigned long long a = 0, b = 0;
    const long long total = (long long)m + 1 + 2 * nnz;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x)
    {
        unsigned long long v;
        if (i <= m) v = (unsigned long long)(unsigned)offs[i];
        else if (i <= m + nnz) v = (unsigned long long)(unsigned)inds[i - m - 1];
        else v = (unsigned long long)__double_as_longlong(vals[i - m - 1 - nnz]);
        const unsigned long long h = mix64(v + 0x9e3779b97f4a7c15ull * (unsigned long long)(i + 1));
        a += h;
        b += mix64(h ^ 0xd6e8feb86659fd93ull);
    }
    for (int o = 16; o > 0; o >>= 1)
    {
        a += __shfl_xor_sync(0xffffffffu, a, o);
        b += __shfl_xor_sync(0xffffffffu, b, o);
    }
    if ((threadIdx.x & 31) == 0)
    {
        atomicAdd(out, a);
        atomicAdd(out + 1, b);
    }
}

// ---- B&B node deltas -------------------------------------------------------------------------------
__global__ void k_node_csr_append(int m0, long long nnz0, int n0, int k, const int *__restrict__ var,
                                  const double *__restrict__ coef, int *__restrict__ offs, int *__restrict__ inds,
                                  double *__restrict__ vals)
{
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= k) return;
    const long long p = nnz0 + 2ll * r;
    inds[p] = var[r];
    vals[p] = coef[r];
    inds[p + 1] = n0 + r;
    vals[p + 1] = -1.0;
    offs[m0 + 1 + r] = (int)(p + 2);
}
// node CSC from the base CSC: column j keeps its base entries (shifted by the branch entries of earlier
// columns) and gains (m0 + r, coef_r) for every branch row on j, appended in row order; the k new slack
// columns follow with one entry (m0 + r, -1) each
__global__ void k_node_csc(int n0, int m0, int k, const int *__restrict__ bptr, const int *__restrict__ brows,
                           const double *__restrict__ bvals, const int *__restrict__ var,
                           const double *__restrict__ coef, int *__restrict__ colptr, int *__restrict__ rows,
                           double *__restrict__ vals)
{
    const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
    for (int j = blockIdx.x * wpb + (threadIdx.x >> 5); j <= n0 + k; j += gridDim.x * wpb)
    {
        if (j >= n0)
        {   // new slack columns and the end pointer
            const int r = j - n0;
            const int start = bptr[n0] + k + r;
            if (lane == 0)
            {
                colptr[j] = start;
                if (r < k)
                {
                    rows[start] = m0 + r;
                    vals[start] = -1.0;
                }
            }
            continue;
        }
        int shift = 0;
        for (int r = 0; r < k; ++r)
            shift += (var[r] < j) ? 1 : 0;
        const int a = bptr[j], e = bptr[j + 1];
        if (lane == 0) colptr[j] = a + shift;
        for (int t = a + lane; t < e; t += 32)
        {
            rows[t + shift] = brows[t];
            vals[t + shift] = bvals[t];
        }
        if (lane == 0)
        {
            int o = e + shift;
            for (int r = 0; r < k; ++r)
                if (var[r] == j)
                {
                    rows[o] = m0 + r;
                    vals[o] = coef[r];
                    ++o;
                }
        }
    }
}
// ---- the deltas of a whole window in ONE launch (grid.y = slot): what the four copies and three kernels per slot of
//      apply_node_delta do, fed from one staged host->device copy ----------------------------------------------------------
struct NodeDeltaDesc
{
    int m0, n0, k, mpad;            // k < 0: the slot keeps its model as it is
    long long nnz0;
    const int *s_var;               // staged on the device: var[k], coef[k], rhs[k]
    const double *s_coef, *s_rhs;
    int *d_var;
    double *d_coef, *b, *c;
    int *csr_offs, *csr_inds;
    double *csr_vals;
    const int *bptr, *brows;
    const double *bvals;
    int *colptr, *rows;
    double *vals;
    double *M;
};
__global__ void k_node_delta_batch(const NodeDeltaDesc *__restrict__ descs)
{
    const NodeDeltaDesc D = descs[blockIdx.y];
    const int k = D.k;
    if (k < 0) return;
    const int m0 = D.m0, n0 = D.n0, m = m0 + k, mpad = D.mpad;
    const int gt = blockIdx.x * blockDim.x + threadIdx.x, gsz = gridDim.x * blockDim.x;
    for (int r = gt; r < k; r += gsz)
    {   // the delta arrays the solver reads, the node's right-hand side and objective tail, its CSR rows
        const int v = D.s_var[r];
        const double cf = D.s_coef[r];
        D.d_var[r] = v;
        D.d_coef[r] = cf;
        D.b[m0 + r] = D.s_rhs[r];
        D.c[n0 + r] = 0.0;
        const long long p = D.nnz0 + 2ll * r;
        D.csr_inds[p] = v;
        D.csr_vals[p] = cf;
        D.csr_inds[p + 1] = n0 + r;
        D.csr_vals[p + 1] = -1.0;
        D.csr_offs[m0 + 1 + r] = (int)(p + 2);
    }
    {   // node CSC from the base CSC (k_node_csc)
        const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
        for (int j = blockIdx.x * wpb + (threadIdx.x >> 5); j <= n0 + k; j += gridDim.x * wpb)
        {
            if (j >= n0)
            {
                const int r = j - n0;
                const int start = D.bptr[n0] + k + r;
                if (lane == 0)
                {
                    D.colptr[j] = start;
                    if (r < k)
                    {
                        D.rows[start] = m0 + r;
                        D.vals[start] = -1.0;
                    }
                }
                continue;
            }
            int shift = 0;
            for (int r = 0; r < k; ++r) shift += (D.s_var[r] < j) ? 1 : 0;
            const int a = D.bptr[j], e = D.bptr[j + 1];
            if (lane == 0) D.colptr[j] = a + shift;
            for (int t = a + lane; t < e; t += 32)
            {
                D.rows[t + shift] = D.brows[t];
                D.vals[t + shift] = D.bvals[t];
            }
            if (lane == 0)
            {
                int o = e + shift;
                for (int r = 0; r < k; ++r)
                    if (D.s_var[r] == j)
                    {
                        D.rows[o] = m0 + r;
                        D.vals[o] = D.s_coef[r];
                        ++o;
                    }
            }
        }
    }
    // identity pad of M: rows m .. mpad-1 whole, columns m .. mpad-1 of the rows above
    const int pad = mpad - m;
    for (long long idx = gt; idx < (long long)pad * mpad; idx += gsz)
    {
        const int r = m + (int)(idx / mpad), c = (int)(idx % mpad);
        D.M[(size_t)r * mpad + c] = (r == c) ? 1.0 : 0.0;
    }
    for (long long idx = gt; idx < (long long)m * pad; idx += gsz)
    {
        const int r = (int)(idx / pad), c = m + (int)(idx % pad);
        D.M[(size_t)r * mpad + c] = 0.0;
    }
}

// rows m0 .. m0+k-1 of M = A D A' for the node (written whole every iteration: the factorisation is in place)
__global__ void k_assemble_extra_rows(int m0, int n0, int k, int ld, const int *__restrict__ var,
                                      const double *__restrict__ coef, const int *__restrict__ bptr,
                                      const int *__restrict__ brows, const double *__restrict__ bvals,
                                      const double *__restrict__ d, double *__restrict__ M)
{
    const int r = blockIdx.x, row = m0 + r, j = var[r];
    const double cf = coef[r], dj = d[j];
    double *Mr = M + (size_t)row * ld;
    for (int c = threadIdx.x; c < row; c += blockDim.x)
        Mr[c] = 0.0;
    __syncthreads();
    for (int t = bptr[j] + threadIdx.x; t < bptr[j + 1]; t += blockDim.x)
        Mr[brows[t]] = cf * bvals[t] * dj;
    for (int q = threadIdx.x; q < r; q += blockDim.x)
        if (var[q] == j) Mr[m0 + q] = cf * coef[q] * dj;
    if (threadIdx.x == 0) Mr[row] = cf * cf * dj + d[n0 + r];
}

// ---- normal-equations operator ---------------------------------------------------------------
void enqueue_factor(sb200_ws *ws, const double *d)
{
    cudaStream_t st = ws->stream;
    if (ws->strategy == SB200_STRATEGY_SYRK)
    {
        launch_syrk_dmma(ws->m, ws->n, ws->denseA, ws->kpad, d, ws->M, ws->mpad, st);
        launch_pad_identity(ws->m, ws->M, ws->mpad, st);
    }
    else
    {
        launch_assemble_normal(ws->pat, d, ws->M, ws->mpad, st);
        if (ws->node_k)
        {
            k_assemble_extra_rows<<<ws->node_k, 256, 0, st>>>(ws->base_m, ws->base_n, ws->node_k, ws->mpad, ws->d_var,
                                                             ws->d_coef, ws->base_colptr, ws->base_rows,
                                                             ws->base_cvals, d, ws->M);
            ++g_launch_count;
        }
    }
    launch_potrf(ws->chol, ws->m, ws->M, ws->mpad, &ws->sc->chol_info, st);
}

// graph-captured chunk of CG iterations
int cg_run(sb200_ws *ws, const double *dscale, double fixed_tol, int cap_override, int honour_done)
{
    cudaStream_t st = ws->stream;
    PcgVecs C{ws->m, ws->V.rhs, ws->cg_diag, ws->cg_x, ws->cg_r, ws->cg_z, ws->cg_p, ws->cg_Ap,
              ws->cg_q, dscale, ws->V.partial};
    launch_cg_init(C, ws->sc, ws->dparams, fixed_tol, honour_done, st);
    const int cap = cap_override > 0 ? cap_override : ws->hparams.cg_max_iter;
    const int chunk = std::max(1, std::min(cap, 32));
    const bool use_graph = ws->params.use_graph != 0;
    if (use_graph && (!ws->cg_graph || ws->cg_graph_chunk != chunk || ws->cg_graph_dscale != dscale))
    {
        if (ws->cg_graph) cudaGraphExecDestroy(ws->cg_graph);
        ws->cg_graph = nullptr;
        cudaGraph_t g = nullptr;
        const long long before = g_launch_count;
        WS_TRY(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
        for (int i = 0; i < chunk; ++i)
            launch_cg_iteration(csr_of(ws), csc_of(ws), C, ws->V, ws->dparams, cap_override, st);
        WS_TRY(cudaStreamEndCapture(st, &g));
        ws->cg_graph_kernels = g_launch_count - before;
        g_launch_count = before;
        WS_TRY(cudaGraphInstantiate(&ws->cg_graph, g, 0));
        cudaGraphDestroy(g);
        ws->cg_graph_chunk = chunk;
        ws->cg_graph_dscale = dscale;
    }
    for (int done_iters = 0; done_iters < cap + chunk; done_iters += chunk)
    {
        if (use_graph)
        {
            WS_TRY(cudaGraphLaunch(ws->cg_graph, st));
            ws->graph_kernel_launches += ws->cg_graph_kernels;
        }
        else
            for (int i = 0; i < chunk; ++i)
                launch_cg_iteration(csr_of(ws), csc_of(ws), C, ws->V, ws->dparams, cap_override, st);
        WS_TRY(cudaMemcpyAsync(ws->sc_host, ws->sc, sizeof(Scalars), cudaMemcpyDeviceToHost, st));
        WS_TRY(cudaStreamSynchronize(st));
        if (ws->sc_host->cg_done || (honour_done && ws->sc_host->done)) break;
    }
    return SB200_OK;
}

// solve (A diag(d) A') z = V.rhs ; result in V.dy
int enqueue_or_run_solve(sb200_ws *ws, const double *dscale, double fixed_tol, int honour_done)
{
    if (ws->strategy == SB200_STRATEGY_PCG)
    {
        int rc = cg_run(ws, dscale, fixed_tol, fixed_tol > 0.0 ? 10000 : 0, honour_done);
        if (rc) return rc;
        if (honour_done) launch_cg_check(ws->sc, ws->stream);
        return SB200_OK;
    }
    launch_potrs(ws->chol, ws->m, ws->M, ws->mpad, ws->V.rhs, ws->stream);
    return SB200_OK;
}

void enqueue_iteration_direct(sb200_ws *ws)
{
    cudaStream_t st = ws->stream;
    const IpmVecs &V = ws->V;
    const CsrView A = csr_of(ws);
    const CscView At = csc_of(ws);
    enqueue_factor(ws, V.d);
    launch_spmv_csr(A, V.t, V.resB, V.rhs, 1.0, 1.0, st);                 // rhs = resB + A t
    launch_potrs(ws->chol, ws->m, ws->M, ws->mpad, V.rhs, st);            // dy (affine)
    launch_spmv_csc(At, CSC_RECOVER, V.dy, nullptr, nullptr, 0, 0, &V, st);
    launch_affine_corrector(V, st);
    launch_spmv_csr(A, V.t, V.resB, V.rhs, 1.0, 1.0, st);
    launch_potrs(ws->chol, ws->m, ws->M, ws->mpad, V.rhs, st);            // dy (corrector)
    launch_spmv_csc(At, CSC_RECOVER, V.dy, nullptr, nullptr, 0, 0, &V, st);
    launch_update(V, ws->dparams, st);
}

int run_iteration_pcg(sb200_ws *ws)
{
    cudaStream_t st = ws->stream;
    const IpmVecs &V = ws->V;
    const CsrView A = csr_of(ws);
    const CscView At = csc_of(ws);
    int rc;
    if (ws->sc_host->done) return SB200_OK;
    launch_jacobi_diag(A, V.d, ws->cg_diag, st);
    launch_spmv_csr(A, V.t, V.resB, V.rhs, 1.0, 1.0, st);
    if ((rc = enqueue_or_run_solve(ws, V.d, 0.0, 1))) return rc;
    launch_spmv_csc(At, CSC_RECOVER, V.dy, nullptr, nullptr, 0, 0, &V, st);
    launch_affine_corrector(V, st);
    launch_spmv_csr(A, V.t, V.resB, V.rhs, 1.0, 1.0, st);
    if ((rc = enqueue_or_run_solve(ws, V.d, 0.0, 1))) return rc;
    launch_spmv_csc(At, CSC_RECOVER, V.dy, nullptr, nullptr, 0, 0, &V, st);
    launch_update(V, ws->dparams, st);
    return SB200_OK;
}

// node = base + delta; no allocation, no device-wide synchronisation once the one-time buffers exist
// one-time / grow-only buffers of the node path: the copy of the base CSC the node CSC is rebuilt from, the delta arrays
static int ensure_node_buffers(sb200_ws *ws, int k, bool sync_copy)
{
    const int n0 = ws->base_n;
    const long long nnz0 = ws->base_nnz;
    cudaStream_t st = ws->stream;
    if (!ws->base_csc_valid)
    {   // one-time copy of the base CSC (the working CSC is rebuilt from it for every node)
        if (ws->node_k != 0) return fail(ws, SB200_ERR_INVALID, "sb200_set_node_delta: base CSC lost");
        int rc;
        if ((rc = grow(ws, &ws->base_colptr, (size_t)n0 + 1))) return rc;
        if ((rc = grow(ws, &ws->base_rows, (size_t)nnz0))) return rc;
        if ((rc = grow(ws, &ws->base_cvals, (size_t)nnz0))) return rc;
        WS_TRY(cudaMemcpyAsync(ws->base_colptr, ws->csc_colptr, sizeof(int) * ((size_t)n0 + 1), cudaMemcpyDeviceToDevice, st));
        WS_TRY(cudaMemcpyAsync(ws->base_rows, ws->csc_rows, sizeof(int) * (size_t)nnz0, cudaMemcpyDeviceToDevice, st));
        WS_TRY(cudaMemcpyAsync(ws->base_cvals, ws->csc_vals, sizeof(double) * (size_t)nnz0, cudaMemcpyDeviceToDevice, st));
        if (sync_copy) WS_TRY(cudaStreamSynchronize(st));     // the batched kernel runs on another stream (the lead's)
        ws->base_csc_valid = true;
    }
    if (k > ws->delta_cap)
    {
        const int cap = (std::max(64, k) + 1) & ~1;     // even: the coefficient staging behind the ids stays 8-byte aligned
        int rc;
        if ((rc = grow(ws, &ws->d_var, (size_t)cap))) return rc;
        if ((rc = grow(ws, &ws->d_coef, (size_t)cap))) return rc;
        if (ws->h_delta) cudaFreeHost(ws->h_delta);
        ws->h_delta = nullptr;
        WS_TRY(cudaMallocHost(&ws->h_delta, (size_t)cap * 20));
        ws->delta_cap = cap;
    }
    return SB200_OK;
}

struct NodeDeltaHost            // a slot's part of a batched delta: the descriptor + the caller's arrays to stage
{
    NodeDeltaDesc d;
    const int *var;
    const double *coef, *rhs;
};

int apply_node_delta(sb200_ws *ws, const sb200_node_delta *delta, NodeDeltaHost *batched = nullptr)
{
    if (batched) batched->d.k = -1;
    if (!ws->loaded) return fail(ws, SB200_ERR_INVALID, "sb200_set_node_delta: no model loaded");
    const int k = delta ? delta->n_extra_rows : 0;
    if (k < 0) return fail(ws, SB200_ERR_INVALID, "sb200_set_node_delta: negative row count");
    ws->warm_ptr = nullptr;
    if (delta && delta->warm_start && delta->warm_n > 0 && delta->warm_m > 0 && delta->warm_n <= ws->base_n_or_n() + k &&
        delta->warm_m <= ws->base_m_or_m() + k)
    {
        ws->warm_ptr = delta->warm_start;
        ws->warm_n = delta->warm_n;
        ws->warm_m = delta->warm_m;
        ws->warm_floor = delta->warm_floor > 0.0 ? delta->warm_floor : 0.1;
    }
    if (k == 0 && ws->node_k == 0) return SB200_OK;
    if (ws->strategy != SB200_STRATEGY_CHOLESKY)
        return fail(ws, SB200_ERR_UNSUPPORTED, "sb200_set_node_delta: only the sparse-assembly + Cholesky strategy folds node rows");
    if (k && (!delta->var || !delta->coef || !delta->rhs))
        return fail(ws, SB200_ERR_INVALID, "sb200_set_node_delta: null delta arrays");
    for (int r = 0; r < k; ++r)       // before any state changes: a rejected delta leaves the workspace as it was
        if (delta->var[r] < 0 || delta->var[r] >= ws->base_n)
            return fail(ws, SB200_ERR_INVALID, "sb200_set_node_delta: branch variable out of range");
    const int m0 = ws->base_m, n0 = ws->base_n;
    const long long nnz0 = ws->base_nnz;
    if (m0 + k > ws->m_cap || n0 + k > ws->n_cap || nnz0 + 2ll * k > ws->nnz_cap ||
        (long long)round_up(m0 + k, SB200_TILE) * round_up(m0 + k, SB200_TILE) > ws->M_cap ||
        round_up(m0 + k, SB200_TILE) / 64 > ws->chol.t_cap)
        return fail(ws, SB200_ERR_UNSUPPORTED,
                    "sb200_set_node_delta: workspace capacity too small (create it with sb200_caps covering the deepest node)");
    WS_TRY(cudaSetDevice(ws->device));
    cudaStream_t st = ws->stream;
    {
        int rc;
        if ((rc = ensure_node_buffers(ws, k, batched != nullptr))) return rc;
    }
    // keep the iteration graph of the depth we leave, pick up the one of the depth we enter
    if (ws->iter_graph) ws->node_graphs[ws->node_k] = std::make_pair(ws->iter_graph, ws->iter_graph_kernels);
    ws->iter_graph = nullptr;
    auto g = ws->node_graphs.find(k);
    if (g != ws->node_graphs.end())
    {
        ws->iter_graph = g->second.first;
        ws->iter_graph_kernels = g->second.second;
    }
    if (batched)
    {   // staged and applied with the rest of the window (apply_node_deltas_batched)
        NodeDeltaDesc &D = batched->d;
        D.m0 = m0; D.n0 = n0; D.k = k; D.mpad = round_up(m0 + k, SB200_TILE);
        D.nnz0 = nnz0;
        D.s_var = nullptr; D.s_coef = nullptr; D.s_rhs = nullptr;
        D.d_var = ws->d_var; D.d_coef = ws->d_coef; D.b = ws->b; D.c = ws->c;
        D.csr_offs = ws->csr_offs; D.csr_inds = ws->csr_inds; D.csr_vals = ws->csr_vals;
        D.bptr = ws->base_colptr; D.brows = ws->base_rows; D.bvals = ws->base_cvals;
        D.colptr = ws->csc_colptr; D.rows = ws->csc_rows; D.vals = ws->csc_vals;
        D.M = ws->M;
        batched->var = k ? delta->var : nullptr;
        batched->coef = k ? delta->coef : nullptr;
        batched->rhs = k ? delta->rhs : nullptr;
    }
    else
    {
        if (k)
        {
            // the previous solve's staging copies have completed (every solve ends with a stream sync)
            int *hv = reinterpret_cast<int *>(ws->h_delta);
            double *hc = reinterpret_cast<double *>(ws->h_delta + 4 * (size_t)ws->delta_cap);
            double *hr = hc + ws->delta_cap;
            for (int r = 0; r < k; ++r)
            {
                if (delta->var[r] < 0 || delta->var[r] >= n0)
                    return fail(ws, SB200_ERR_INVALID, "sb200_set_node_delta: branch variable out of range");
                hv[r] = delta->var[r];
                hc[r] = delta->coef[r];
                hr[r] = delta->rhs[r];
            }
            WS_TRY(cudaMemcpyAsync(ws->d_var, hv, sizeof(int) * k, cudaMemcpyHostToDevice, st));
            WS_TRY(cudaMemcpyAsync(ws->d_coef, hc, sizeof(double) * k, cudaMemcpyHostToDevice, st));
            WS_TRY(cudaMemcpyAsync(ws->b + m0, hr, sizeof(double) * k, cudaMemcpyHostToDevice, st));
            WS_TRY(cudaMemsetAsync(ws->c + n0, 0, sizeof(double) * k, st));
            k_node_csr_append<<<(k + 63) / 64, 64, 0, st>>>(m0, nnz0, n0, k, ws->d_var, ws->d_coef, ws->csr_offs,
                                                           ws->csr_inds, ws->csr_vals);
            ++g_launch_count;
        }
        k_node_csc<<<grid_for((long long)(n0 + k + 1) * 32, 256, 148 * 8), 256, 0, st>>>(
            n0, m0, k, ws->base_colptr, ws->base_rows, ws->base_cvals, ws->d_var, ws->d_coef, ws->csc_colptr, ws->csc_rows,
            ws->csc_vals);
        ++g_launch_count;
    }
    ws->m = m0 + k;
    ws->n = n0 + k;
    ws->nnz = nnz0 + 2ll * k;
    ws->mpad = round_up(ws->m, SB200_TILE);
    ws->node_k = k;
    if (!batched) launch_pad_identity(ws->m, ws->M, ws->mpad, st);
    int rc = chol_work_ensure(ws->err, ws->chol, ws->mpad);
    if (rc) return rc;
    carve(ws);
    return SB200_OK;
}

// Capture one loop iteration.  If the driver refuses to capture (e.g. cooperative launches inside a
// capture), fall back to plain stream launches for this workspace - same kernels, more launch gaps.
int ensure_iter_graph(sb200_ws *ws)
{
    if (ws->iter_graph) return SB200_OK;
    cudaGraph_t g = nullptr;
    const long long before = g_launch_count;
    cudaError_t e = cudaStreamBeginCapture(ws->stream, cudaStreamCaptureModeThreadLocal);
    if (e == cudaSuccess)
    {
        enqueue_iteration_direct(ws);
        e = cudaStreamEndCapture(ws->stream, &g);
    }
    ws->iter_graph_kernels = g_launch_count - before;
    g_launch_count = before;
    if (e == cudaSuccess && ws->node_graphs.size() >= 2)
    {   // B&B: a new depth has the same graph topology with other kernel arguments.  Instantiating costs
        // milliseconds per workspace (measured: 4-5 ms, a 100 ms stall of a 16-slot window whenever the search
        // reaches a new level); updating the executable graph of the shallowest cached depth in place does not.
        auto victim = ws->node_graphs.begin();
        cudaGraphExecUpdateResultInfo info{};
        if (victim->second.first && cudaGraphExecUpdate(victim->second.first, g, &info) == cudaSuccess)
        {
            ws->iter_graph = victim->second.first;
            ws->node_graphs.erase(victim);
        }
        else
            cudaGetLastError();
    }
    if (e == cudaSuccess && !ws->iter_graph) e = cudaGraphInstantiate(&ws->iter_graph, g, 0);
    if (g) cudaGraphDestroy(g);
    if (e != cudaSuccess)
    {
        cudaGetLastError();
        ws->iter_graph = nullptr;
        ws->params.use_graph = 0;
        ws->err.msg = std::string("graph capture unavailable, using stream launches: ") + cudaGetErrorString(e);
    }
    return SB200_OK;
}

// the one-CTA solver needs the compact symbolic structure (unit products, 2-byte ids), the sparse-assembly +
// Cholesky strategy and a matrix whose solve vectors fit shared memory
bool cta_eligible(const sb200_ws *ws)
{
    return ws->strategy == SB200_STRATEGY_CHOLESKY && ws->pat.term16 && ws->pat.chunk_ptr && ws->pat.pad_id >= 0 &&
           ws->mpad <= CTA_MAX_MPAD && ws->chol.linv && ws->csr_offs && ws->csc_colptr;
}

// everything k_ipm_cta needs for the model as it stands in the workspace (base model + applied node delta)
void fill_cta_args(sb200_ws *ws, const sb200_result *res, CtaLp &c)
{
    c.V = ws->V;
    c.P = ws->dparams;
    c.csr_offs = ws->csr_offs; c.csr_inds = ws->csr_inds; c.csr_vals = ws->csr_vals;
    c.csc_colptr = ws->csc_colptr; c.csc_rows = ws->csc_rows; c.csc_vals = ws->csc_vals;
    c.n_pairs = ws->pat.n_pairs;
    c.chunk_ptr = ws->pat.chunk_ptr;
    c.term8 = reinterpret_cast<const uint4 *>(ws->pat.term16);
    c.nd = ws->pat.pad_id + 1;
    c.ones = ws->ones_n;
    c.base_m = ws->node_k ? ws->base_m : ws->m;
    c.base_n = ws->node_k ? ws->base_n : ws->n;
    c.node_k = ws->node_k;
    c.d_var = ws->d_var; c.d_coef = ws->d_coef;
    c.base_colptr = ws->base_colptr; c.base_rows = ws->base_rows; c.base_cvals = ws->base_cvals;
    c.M = ws->M;
    c.ld = ws->mpad;
    c.linv = ws->chol.linv;
    {   // pattern-only products when the lists exist and their staging fits the block's shared memory
        const int bm = c.base_m, bn = c.base_n, k = c.node_k;
        const bool fit = ws->lists.row16 && ws->lists.m == bm && ws->lists.n == bn && cta_lists_fit(bm, bn, k);
        c.row_ptr = fit ? ws->lists.row_ptr : nullptr;
        c.col_ptr = fit ? ws->lists.col_ptr : nullptr;
        c.row16 = fit ? reinterpret_cast<const uint4 *>(ws->lists.row16) : nullptr;
        c.col16 = fit ? reinterpret_cast<const uint4 *>(ws->lists.col16) : nullptr;
        c.col_sign = fit ? ws->lists.col_sign : nullptr;
    }
    c.warm = ws->warm_ptr;
    c.warm_n = ws->warm_n;
    c.warm_m = ws->warm_m;
    c.warm_floor = ws->warm_floor;
    ws->warm_ptr = nullptr;                    // one use
    c.export_xys = res ? res->xys_device : nullptr;
    c.sc_pinned = ws->sc_host;
}

// ---- solve state machine ------------------------------------------------------------------------
int solve_begin(sb200_ws *ws, const sb200_params *p, sb200_result *res)
{
    if (!ws->loaded) return fail(ws, SB200_ERR_INVALID, "sb200_solve: no model loaded");
    WS_TRY(cudaSetDevice(ws->device));
    ws->params = *p;
    if (ws->params.poll_every < 1) ws->params.poll_every = 1;
    cudaStream_t st = ws->stream;
    const IpmVecs &V = ws->V;
    const CsrView A = csr_of(ws);
    const CscView At = csc_of(ws);
    ws->launches_at_begin = g_launch_count;
    ws->graph_kernel_launches = 0;

    DevParams &hp = ws->hparams;
    hp.eta = p->eta;
    hp.mu_tol = p->mu_tol;
    hp.min_improv_ratio = p->gap_min_improv_pct / 100.0;
    hp.max_iter = p->max_iter;
    hp.gap_enabled = (p->gap_enabled && p->gap_window > 0 && p->gap_min_improv_pct >= 0.0) ? 1 : 0;
    hp.gap_window = p->gap_window;
    hp.n_orig = ws->n_orig;
    hp.cg_max_iter = p->cg_max_iter;
    hp.cg_tol_initial = p->cg_tol_initial;
    hp.cg_tol_final = p->cg_tol_final;
    hp.cg_tol_decay = p->cg_tol_decay;
    WS_TRY(cudaMemcpyAsync(ws->dparams, &hp, sizeof hp, cudaMemcpyHostToDevice, st));

    WS_TRY(cudaEventRecord(ws->ev[0], st));
    launch_reset_scalars(ws->sc, st);

    ws->cta_launched = false;
    if (ws->solver_form == SB200_FORM_THROUGHPUT && cta_eligible(ws) &&
        !(res && (res->x0_host || res->y0_host || res->s0_host)))
    {   // one launch, one CTA: starting point, loop and termination test on the device (sb200_cta.cu)
        if (!ws->cta_dev)
        {
            WS_TRY(cudaMalloc(&ws->cta_dev, sizeof(CtaLp)));
            WS_TRY(cudaMallocHost(&ws->cta_host, sizeof(CtaLp)));
        }
        fill_cta_args(ws, res, *ws->cta_host);
        WS_TRY(cudaMemcpyAsync(ws->cta_dev, ws->cta_host, sizeof(CtaLp), cudaMemcpyHostToDevice, st));
        WS_TRY(cudaEventRecord(ws->ev[1], st));
        WS_TRY(cudaEventRecord(ws->ev[2], st));
        const int rc = launch_ipm_cta(ws->cta_dev, 1, st);
        if (rc) return fail(ws, rc, "sb200_solve: launch of the one-CTA solver failed");
        WS_TRY(cudaGetLastError());
        ws->cta_launched = true;
        ws->enqueued = ws->params.max_iter;        // nothing left to enqueue: solve_step only requests the scalar block
        ws->active = true;
        ws->sc_host->done = 0;
        return SB200_OK;
    }
    ws->warm_ptr = nullptr;                        // the latency form starts cold

    // ---- starting point (sypha_solver_init.cpp:543-652), D = I ---------------------------------
    int rc;
    WS_TRY(cudaMemsetAsync(V.rhs, 0, sizeof(double) * ws->mpad, st));
    if (ws->strategy == SB200_STRATEGY_PCG)
        launch_jacobi_diag(A, ws->ones_n, ws->cg_diag, st);
    else
        enqueue_factor(ws, ws->ones_n);
    WS_TRY(cudaMemcpyAsync(V.rhs, ws->b, sizeof(double) * ws->m, cudaMemcpyDeviceToDevice, st));
    if ((rc = enqueue_or_run_solve(ws, nullptr, 1e-12, 0))) return rc;        // (AA')^-1 b
    launch_spmv_csc(At, CSC_START_X, V.dy, nullptr, nullptr, 0, 0, &V, st);  // x~ = A' (.)
    launch_spmv_csr(A, ws->c, nullptr, V.rhs, 1.0, 0.0, st);                 // A c
    if ((rc = enqueue_or_run_solve(ws, nullptr, 1e-12, 0))) return rc;        // y~
    WS_TRY(cudaMemcpyAsync(V.y, V.dy, sizeof(double) * ws->m, cudaMemcpyDeviceToDevice, st));
    launch_spmv_csc(At, CSC_START_S, V.y, nullptr, nullptr, 0, 0, &V, st);   // s~ = c - A' y~
    launch_start_shift1(V, st);
    launch_start_shift2(V, st);
    if (res)
    {   // node.hX / hY / hS = starting point (sypha_solver.cpp:72-78)
        if (res->x0_host) WS_TRY(cudaMemcpyAsync(res->x0_host, V.x, sizeof(double) * ws->n, cudaMemcpyDeviceToHost, st));
        if (res->y0_host) WS_TRY(cudaMemcpyAsync(res->y0_host, V.y, sizeof(double) * ws->m, cudaMemcpyDeviceToHost, st));
        if (res->s0_host) WS_TRY(cudaMemcpyAsync(res->s0_host, V.s, sizeof(double) * ws->n, cudaMemcpyDeviceToHost, st));
    }
    WS_TRY(cudaEventRecord(ws->ev[1], st));

    // ---- initial residuals and mu (sypha_solver.cpp:375-459) ------------------------------------
    launch_spmv_csc(At, CSC_RESC, V.y, nullptr, nullptr, 0, 0, &V, st);      // resC = c - s - A'y
    launch_spmv_csr(A, V.x, ws->b, V.resB, -1.0, 1.0, st);                   // resB = b - A x
    launch_init_mu(V, ws->dparams, st);
    launch_prologue(V, st);
    WS_TRY(cudaEventRecord(ws->ev[2], st));
    ws->enqueued = 0;
    ws->active = true;
    ws->sc_host->done = 0;
    return SB200_OK;
}

// enqueue up to poll_every iterations and the scalar-block read-back
int solve_step(sb200_ws *ws)
{
    cudaStream_t st = ws->stream;
    if (ws->cta_launched)
    {   // the kernel mirrors the scalar block into pinned memory itself; the event is for callers that block
        WS_TRY(cudaEventRecord(ws->ev[3], st));
        return SB200_OK;
    }
    const int room = ws->params.max_iter - ws->enqueued;
    const int k = std::min(ws->params.poll_every, std::max(room, 0));
    for (int i = 0; i < k; ++i)
    {
        if (ws->strategy == SB200_STRATEGY_PCG)
        {
            int rc = run_iteration_pcg(ws);
            if (rc) return rc;
        }
        else if (ws->params.use_graph)
        {
            int rc = ensure_iter_graph(ws);
            if (rc) return rc;
            if (ws->iter_graph)
            {
                WS_TRY(cudaGraphLaunch(ws->iter_graph, st));
                ws->graph_kernel_launches += ws->iter_graph_kernels;
            }
            else
                enqueue_iteration_direct(ws);
        }
        else
            enqueue_iteration_direct(ws);
    }
    ws->enqueued += k;
    WS_TRY(cudaMemcpyAsync(ws->sc_host, ws->sc, sizeof(Scalars), cudaMemcpyDeviceToHost, st));
    WS_TRY(cudaEventRecord(ws->ev[3], st));
    return SB200_OK;
}

// wait for the last step; returns 1 when the loop has finished
int solve_poll(sb200_ws *ws, int *finished)
{
    WS_TRY(cudaEventSynchronize(ws->ev[3]));
    const bool stop = (ws->params.stop_flag && *ws->params.stop_flag) ||
                      (ws->params.stop_cb && ws->params.stop_cb(ws->params.stop_user));
    *finished = (ws->sc_host->done || ws->enqueued >= ws->params.max_iter || stop) ? 1 : 0;
    return SB200_OK;
}

// result of a one-block solve from the pinned mirror the kernel filled (no copy, no synchronisation here: the caller has
// seen the LP finish - an event, a stream synchronisation or the mirror's `done`)
void finish_from_mirror(sb200_ws *ws, sb200_result *r)
{
    const Scalars &sc = *ws->sc_host;
    ws->active = false;
    const bool numerical = sc.numerical != 0 || sc.chol_info != 0;
    int reason = sc.reason;
    if (numerical) reason = SB200_TERM_INFEASIBLE_OR_NUMERICAL;
    const double viol = sc.dual - sc.primal;
    bool num2 = numerical;
    if (!numerical && reason != SB200_TERM_CONVERGED && std::isfinite(viol) && viol > 1e6 * std::max(1.0, std::fabs(sc.primal)))
    {
        num2 = true;
        reason = SB200_TERM_INFEASIBLE_OR_NUMERICAL;
    }
    r->status = num2 ? SB200_ERR_NUMERICAL : SB200_OK;
    r->reason = reason;
    r->iterations = sc.iter;
    r->primal_obj = sc.primal;
    r->dual_obj = sc.dual;
    r->rel_gap = std::fabs(sc.primal - sc.dual) / std::max(1.0, std::fabs(sc.primal));
    r->mu = sc.mu;
    r->strategy_used = ws->strategy;
    r->cg_iterations = 0;
    // device time of the LP: the kernel's own clock (%globaltimer around the whole solve, mirrored in sum0)
    r->ms_start = 0.0;
    r->ms_setup = 0.0;
    r->ms_loop = sc.sum0 * 1e-6;
    r->kernels_launched = (g_launch_count - ws->launches_at_begin) + ws->graph_kernel_launches;
    ws->trace_rows = std::min(sc.iter, SB200_TRACE_ROWS);
    ws->trace_stale = true;                 // copied on demand (sb200_get_trace)
}

int solve_finish(sb200_ws *ws, sb200_result *r)
{
    cudaStream_t st = ws->stream;
    const IpmVecs &V = ws->V;
    if (ws->cta_launched && !r->x_host && !r->y_host && !r->s_host)
    {   // one-block solve, nothing to copy out: the kernel left the scalars in the pinned mirror (and the packed iterate
        // in xys_device) and the caller has seen it finish
        finish_from_mirror(ws, r);
        return SB200_OK;
    }
    if (ws->cta_launched && r->xys_device) r->xys_device = nullptr;      // already written by the kernel
    WS_TRY(cudaMemcpyAsync(ws->sc_host, ws->sc, sizeof(Scalars), cudaMemcpyDeviceToHost, st));
    if (r->x_host) WS_TRY(cudaMemcpyAsync(r->x_host, V.x, sizeof(double) * ws->n, cudaMemcpyDeviceToHost, st));
    if (r->y_host) WS_TRY(cudaMemcpyAsync(r->y_host, V.y, sizeof(double) * ws->m, cudaMemcpyDeviceToHost, st));
    if (r->s_host) WS_TRY(cudaMemcpyAsync(r->s_host, V.s, sizeof(double) * ws->n, cudaMemcpyDeviceToHost, st));
    if (r->xys_device)
    {   // the iterate a child node may start from: x | y | s packed, device to device, before the final sync
        WS_TRY(cudaMemcpyAsync(r->xys_device, V.x, sizeof(double) * ws->n, cudaMemcpyDeviceToDevice, st));
        WS_TRY(cudaMemcpyAsync(r->xys_device + ws->n, V.y, sizeof(double) * ws->m, cudaMemcpyDeviceToDevice, st));
        WS_TRY(cudaMemcpyAsync(r->xys_device + ws->n + ws->m, V.s, sizeof(double) * ws->n, cudaMemcpyDeviceToDevice, st));
    }
    WS_TRY(cudaStreamSynchronize(st));
    const Scalars &sc = *ws->sc_host;
    ws->active = false;

    int reason = sc.reason;
    int trsv_err = 0;
    if (ws->chol.ctl) WS_TRY(cudaMemcpy(&trsv_err, ws->chol.ctl + 8, sizeof(int), cudaMemcpyDeviceToHost));
    if (trsv_err)
    {
        cudaMemset(ws->chol.ctl + 8, 0, sizeof(int));
        return fail(ws, SB200_ERR_CUDA, "data-flow wait timed out in the factorisation / triangular solves");
    }
    bool numerical = sc.numerical != 0 || sc.chol_info != 0;
    // sypha_solver.cpp:775-778: anything but a failure / gap stall is re-labelled from mu
    // (this also overwrites TIME_LIMIT, as the reference does)
    if (!numerical && reason != SB200_TERM_GAP_STALLED)
        reason = (sc.mu <= ws->params.mu_tol) ? SB200_TERM_CONVERGED : SB200_TERM_MAX_ITER;
    if (numerical) reason = SB200_TERM_INFEASIBLE_OR_NUMERICAL;
    const double viol = sc.dual - sc.primal;
    if (!numerical && reason != SB200_TERM_CONVERGED && std::isfinite(viol) &&
        viol > 1e6 * std::max(1.0, std::fabs(sc.primal)))      // :788-796
    {
        numerical = true;
        reason = SB200_TERM_INFEASIBLE_OR_NUMERICAL;
    }
    r->status = numerical ? SB200_ERR_NUMERICAL : SB200_OK;
    r->reason = reason;
    r->iterations = sc.iter;
    r->primal_obj = sc.primal;
    r->dual_obj = sc.dual;
    r->rel_gap = std::fabs(sc.primal - sc.dual) / std::max(1.0, std::fabs(sc.primal));
    r->mu = sc.mu;
    r->strategy_used = ws->strategy;
    r->cg_iterations = sc.cg_total;
    float ms = 0.f;
    cudaEventElapsedTime(&ms, ws->ev[0], ws->ev[1]); r->ms_start = ms;
    cudaEventElapsedTime(&ms, ws->ev[1], ws->ev[2]); r->ms_setup = ms;
    cudaEventElapsedTime(&ms, ws->ev[2], ws->ev[3]); r->ms_loop = ms;
    r->kernels_launched = (g_launch_count - ws->launches_at_begin) + ws->graph_kernel_launches;
    ws->trace_rows = std::min(sc.iter, SB200_TRACE_ROWS);
    ws->trace_stale = false;
    ws->trace_host.resize((size_t)ws->trace_rows * SB200_TRACE_COLS);
    if (ws->trace_rows)
        WS_TRY(cudaMemcpy(ws->trace_host.data(), V.trace, sizeof(double) * ws->trace_host.size(),
                          cudaMemcpyDeviceToHost));
    return SB200_OK;
}

} // namespace

// =====================================================================================================
// C ABI
// =====================================================================================================
extern "C" {

int sb200_version(void) { return SB200_VERSION; }

int sb200_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess)
    {
        cudaGetLastError();
        return 0;
    }
    return n;
}

void sb200_default_params(sb200_params *p)
{
    memset(p, 0, sizeof *p);
    p->max_iter = 25;          // kMehrotraMaxIter
    p->eta = 0.95;             // kMehrotraEta
    p->mu_tol = 1e-4;          // kMehrotraMuTol
    p->gap_enabled = 0;
    p->strategy = SB200_STRATEGY_AUTO;
    p->cg_max_iter = 500;
    p->cg_tol_initial = 1e-2;
    p->cg_tol_final = 1e-8;
    p->cg_tol_decay = 0.5;
    p->stop_flag = nullptr;
    p->stop_cb = nullptr;
    p->stop_user = nullptr;
    p->poll_every = 1;
    p->use_graph = 1;
}

int sb200_ws_create(int device, const sb200_caps *caps, sb200_ws **out)
{
    if (!out) return SB200_ERR_INVALID;
    *out = nullptr;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
    {
        cudaGetLastError();
        return SB200_ERR_CUDA;     // no CPU fallback: the library needs a GPU
    }
    if (device < 0) device = 0;
    if (device >= ndev) return SB200_ERR_INVALID;
    sb200_ws *ws = new sb200_ws();
    ws->device = device;
    auto bail = [&](int rc) { sb200_ws_destroy(ws); return rc; };
    if (cudaSetDevice(device) != cudaSuccess) return bail(SB200_ERR_CUDA);
    if (cudaStreamCreateWithFlags(&ws->stream, cudaStreamNonBlocking) != cudaSuccess) return bail(SB200_ERR_CUDA);
    {   // set-up temporaries and symbolic structures come from the stream-ordered pool; keep what it has
        // freed cached, so that re-loading a model (every e2e step, every base-model change of the B&B
        // driver) does not go back to the driver for memory
        cudaMemPool_t pool = nullptr;
        if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess && pool)
        {
            unsigned long long keep = ~0ull;
            cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
        }
        cudaGetLastError();
    }
    for (auto &e : ws->ev)
        if (cudaEventCreate(&e) != cudaSuccess) return bail(SB200_ERR_CUDA);
    if (cudaMalloc(&ws->sc, sizeof(Scalars)) != cudaSuccess) return bail(SB200_ERR_NOMEM);
    if (cudaMalloc(&ws->dparams, sizeof(DevParams)) != cudaSuccess) return bail(SB200_ERR_NOMEM);
    if (cudaMallocHost(&ws->sc_host, sizeof(Scalars)) != cudaSuccess) return bail(SB200_ERR_NOMEM);
    memset(ws->sc_host, 0, sizeof(Scalars));
    if (caps && caps->m_max > 0 && caps->n_max > 0 && caps->nnz_max > 0)
    {
        int rc = ensure_capacity(ws, caps->m_max, caps->n_max, caps->nnz_max);
        if (rc) return bail(rc);
    }
    *out = ws;
    return SB200_OK;
}

int sb200_ws_destroy(sb200_ws *ws)
{
    if (!ws) return SB200_OK;
    cudaSetDevice(ws->device);
    if (ws->stream) cudaStreamSynchronize(ws->stream);
    drop_graphs(ws);
    free_normal_pattern(&ws->pat);
    free_compact_lists(&ws->lists);
    free_blocked(&ws->blk_rows);
    free_blocked(&ws->blk_cols);
    chol_work_free(ws->chol);
    void *ptrs[] = {ws->csr_offs, ws->csr_inds, ws->csr_vals, ws->csc_colptr, ws->csc_rows, ws->csc_vals,
                    ws->c, ws->b, ws->denseA, ws->M, ws->slab, ws->sc, ws->dparams, ws->base_colptr, ws->base_rows,
                    ws->base_cvals, ws->fp_dev, ws->d_var, ws->d_coef, ws->heur_list, ws->heur_sorted, ws->heur_cover, ws->heur_nif, ws->heur_score, ws->heur_out};
    if (ws->h_delta) cudaFreeHost(ws->h_delta);
    if (ws->nd_host) cudaFreeHost(ws->nd_host);
    if (ws->nd_dev) cudaFree(ws->nd_dev);
    if (ws->heur_out_host) cudaFreeHost(ws->heur_out_host);
    if (ws->heur_flag_host) cudaFreeHost(ws->heur_flag_host);
    if (ws->cta_dev) cudaFree(ws->cta_dev);
    if (ws->cta_host) cudaFreeHost(ws->cta_host);
    if (ws->batch_dev) cudaFree(ws->batch_dev);
    if (ws->batch_host) cudaFreeHost(ws->batch_host);
    if (ws->hbatch_dev) cudaFree(ws->hbatch_dev);
    if (ws->hbatch_host) cudaFreeHost(ws->hbatch_host);
    for (void *p : ptrs)
        if (p) cudaFree(p);
    if (ws->sc_host) cudaFreeHost(ws->sc_host);
    for (auto &e : ws->ev)
        if (e) cudaEventDestroy(e);
    if (ws->stream) cudaStreamDestroy(ws->stream);
    delete ws;
    return SB200_OK;
}

const char *sb200_last_error(const sb200_ws *ws) { return ws ? ws->err.msg.c_str() : "null workspace"; }

void *sb200_stream(sb200_ws *ws) { return ws ? (void *)ws->stream : nullptr; }

int sb200_load_model(sb200_ws *ws, int m, int n, int n_orig, long long nnz, const int *csr_offs,
                     const int *csr_inds, const double *csr_vals, const double *c, const double *b,
                     int ptrs_on_device, int strategy_hint)
{
    if (!ws) return SB200_ERR_INVALID;
    if (m <= 0 || n <= 0 || nnz <= 0 || n_orig < 0 || n_orig > n || !csr_offs || !csr_inds || !csr_vals ||
        !c || !b || nnz > 0x7fffffffll)
        return fail(ws, SB200_ERR_INVALID, "sb200_load_model: bad dimensions or null pointer");
    WS_TRY(cudaSetDevice(ws->device));
    cudaStream_t st = ws->stream;
    const bool same_shape = ws->loaded && ws->fp_valid && ws->node_k == 0 && ws->m == m && ws->n == n && ws->n_orig == n_orig &&
                            ws->nnz == nnz && ws->fp_strategy_hint == strategy_hint && ws->strategy != SB200_STRATEGY_SYRK;
    if (same_shape)
    {   // candidate for the fast path: copy the arrays in, fingerprint them, compare
        const cudaMemcpyKind kd = ptrs_on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
        WS_TRY(cudaMemcpyAsync(ws->csr_offs, csr_offs, sizeof(int) * ((size_t)m + 1), kd, st));
        WS_TRY(cudaMemcpyAsync(ws->csr_inds, csr_inds, sizeof(int) * (size_t)nnz, kd, st));
        WS_TRY(cudaMemcpyAsync(ws->csr_vals, csr_vals, sizeof(double) * (size_t)nnz, kd, st));
        WS_TRY(cudaMemcpyAsync(ws->c, c, sizeof(double) * (size_t)n, kd, st));
        WS_TRY(cudaMemcpyAsync(ws->b, b, sizeof(double) * (size_t)m, kd, st));
        unsigned long long fpn[2];
        WS_TRY(cudaMemsetAsync(ws->fp_dev, 0, 2 * sizeof(unsigned long long), st));
        k_fingerprint<<<grid_for((long long)m + 1 + 2 * nnz, 256, 148 * 8), 256, 0, st>>>(m, nnz, ws->csr_offs, ws->csr_inds,
                                                                                          ws->csr_vals, ws->fp_dev);
        ++g_launch_count;
        WS_TRY(cudaMemcpyAsync(fpn, ws->fp_dev, sizeof fpn, cudaMemcpyDeviceToHost, st));
        WS_TRY(cudaStreamSynchronize(st));
        if (fpn[0] == ws->fp[0] && fpn[1] == ws->fp[1])
        {   // the same matrix as the resident one: CSC, symbolic structure, blocked copies, M, the factorisation's task
            // lists and the captured iteration graph all stay; iterates are overwritten by the next solve's start
            ws->loaded = true;
            return SB200_OK;
        }
    }
    ws->loaded = false;
    ws->fp_valid = false;
    drop_graphs(ws);
    int rc = ensure_capacity(ws, m, n, nnz);
    if (rc) return rc;
    ws->m = m; ws->n = n; ws->n_orig = n_orig; ws->nnz = nnz;
    ws->base_m = m; ws->base_n = n; ws->base_nnz = nnz;
    ws->node_k = 0;
    ws->base_csc_valid = false;
    ws->mpad = round_up(m, SB200_TILE);
    const cudaMemcpyKind kind = ptrs_on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
    WS_TRY(cudaMemcpyAsync(ws->csr_offs, csr_offs, sizeof(int) * ((size_t)m + 1), kind, st));
    WS_TRY(cudaMemcpyAsync(ws->csr_inds, csr_inds, sizeof(int) * (size_t)nnz, kind, st));
    WS_TRY(cudaMemcpyAsync(ws->csr_vals, csr_vals, sizeof(double) * (size_t)nnz, kind, st));
    WS_TRY(cudaMemcpyAsync(ws->c, c, sizeof(double) * (size_t)n, kind, st));
    WS_TRY(cudaMemcpyAsync(ws->b, b, sizeof(double) * (size_t)m, kind, st));
    rc = build_csc(ws->err, m, n, nnz, ws->csr_offs, ws->csr_inds, ws->csr_vals, ws->csc_colptr,
                   ws->csc_rows, ws->csc_vals, st);
    if (rc) return rc;
    ws->csc_lanes = pick_csc_lanes(nnz, n);
    {
        if (!ws->fp_dev) WS_TRY(cudaMalloc(&ws->fp_dev, 2 * sizeof(unsigned long long)));
        int *flag = reinterpret_cast<int *>(ws->fp_dev), dup = 0;
        WS_TRY(cudaMemsetAsync(flag, 0, sizeof(int), st));
        k_find_duplicates<<<grid_for(n, 256, 148 * 8), 256, 0, st>>>(n, ws->csc_colptr, ws->csc_rows, flag);
        ++g_launch_count;
        WS_TRY(cudaMemcpyAsync(&dup, flag, sizeof(int), cudaMemcpyDeviceToHost, st));
        WS_TRY(cudaStreamSynchronize(st));
        if (dup) return fail(ws, SB200_ERR_INVALID, "sb200_load_model: a (row, column) pair is stored more than once");
    }

    // ---- strategy --------------------------------------------------------------------------------
    int strat = strategy_hint;
    if (strat == SB200_STRATEGY_AUTO)
        strat = (m <= 16384) ? SB200_STRATEGY_CHOLESKY : SB200_STRATEGY_PCG;
    free_normal_pattern(&ws->pat, st);
    if (strat == SB200_STRATEGY_CHOLESKY)
    {
        // pad id of the compact term lists = n_cap: never a real column, d[n_cap] = 0 (also for B&B nodes
        // whose extra slack columns take the ids n .. n_cap-1)
        rc = build_normal_pattern(ws->err, m, n, nnz, ws->csc_colptr, ws->csc_rows, ws->csc_vals, &ws->pat, st, ws->n_cap);
        if (rc == SB200_ERR_UNSUPPORTED && strategy_hint == SB200_STRATEGY_AUTO)
            strat = SB200_STRATEGY_PCG;
        else if (rc)
            return rc;
        else if (strategy_hint == SB200_STRATEGY_AUTO)
        {
            // density switch (north_star item 1): sparse gather moves 4..12 B per product term at
            // HBM speed, SYRK spends m^2 n flops on the FP64 tensor pipe.
            const double t_sparse = (double)ws->pat.n_terms * (ws->pat.term_w ? 12.0 : 4.0) / 6.5e12;
            const double t_syrk = (double)m * (double)m * (double)n / 3.0e13;
            if (t_syrk < t_sparse && (double)ws->mpad * round_up(n, 32) * 8.0 < 40e9)
            {
                strat = SB200_STRATEGY_SYRK;
                free_normal_pattern(&ws->pat, st);
            }
        }
    }
    ws->strategy = strat;
    free_compact_lists(&ws->lists, st);
    if (strat == SB200_STRATEGY_CHOLESKY && ws->pat.term16 && ws->mpad <= CTA_MAX_MPAD)
    {   // the one-block solver's products read 2-byte pattern lists instead of the 12-byte CSR / CSC entries
        const char *off = getenv("SB200_COMPACT_PRODUCTS");
        if (!(off && off[0] == '0'))
        {
            rc = build_compact_lists(ws->err, m, n, ws->csr_offs, ws->csr_inds, ws->csc_colptr, ws->csc_rows, ws->csc_vals,
                                     &ws->lists, st);
            if (rc != SB200_OK && rc != SB200_ERR_UNSUPPORTED) return rc;
        }
    }
    free_blocked(&ws->blk_rows, st);
    free_blocked(&ws->blk_cols, st);
    if (strat == SB200_STRATEGY_PCG)
    {   // +/-1 matrix => pattern-only, shared-memory-staged products (sb200_blocked.cu)
        const char *off = getenv("SB200_BLOCKED");
        if (!(off && off[0] == '0'))
        {
            rc = build_blocked(ws->err, m, n, nnz, ws->csr_offs, ws->csr_inds, ws->csr_vals, blocked_nb_for_rows(n), true,
                               &ws->blk_rows, st);
            if (rc == SB200_OK)
                rc = build_blocked(ws->err, n, m, nnz, ws->csc_colptr, ws->csc_rows, ws->csc_vals, blocked_nb_for_cols(m),
                                   false, &ws->blk_cols, st);
            if (rc != SB200_OK)
            {
                free_blocked(&ws->blk_rows);
                free_blocked(&ws->blk_cols);
                if (rc != SB200_ERR_UNSUPPORTED) return rc;      // general coefficients: value-carrying kernels
            }
        }
    }
    if (strat == SB200_STRATEGY_SYRK)
    {
        ws->kpad = round_up(n, 32);
        if (ws->denseA) cudaFree(ws->denseA);
        ws->denseA = nullptr;
        const size_t bytes = sizeof(double) * (size_t)ws->mpad * ws->kpad;
        WS_TRY(cudaMalloc(&ws->denseA, bytes));
        WS_TRY(cudaMemsetAsync(ws->denseA, 0, bytes, st));
        k_densify<<<grid_for((long long)m * 32, 256, 148 * 16), 256, 0, st>>>(m, ws->csr_offs, ws->csr_inds,
                                                                             ws->csr_vals, ws->denseA, ws->kpad);
        ++g_launch_count;
    }
    if (strat != SB200_STRATEGY_PCG)
    {
        const int mpad_cap = round_up(ws->m_cap, SB200_TILE);      // deepest B&B node the workspace was sized for
        const long long need = (long long)mpad_cap * mpad_cap;
        if (need > ws->M_cap)
        {
            if (ws->M) cudaFree(ws->M);
            ws->M = nullptr;
            WS_TRY(cudaMalloc(&ws->M, sizeof(double) * (size_t)need));
            ws->M_cap = need;
        }
        WS_TRY(cudaMemsetAsync(ws->M, 0, sizeof(double) * (size_t)need, st));
        launch_pad_identity(m, ws->M, ws->mpad, st);
        rc = chol_work_ensure(ws->err, ws->chol, ws->mpad, mpad_cap);
        if (rc) return rc;
    }
    carve(ws);
    WS_TRY(cudaMemsetAsync(ws->slab, 0, ws->slab_bytes, st));
    launch_fill(ws->ones_n, 1.0, ws->n_cap, st);   // entry n_cap stays 0: pad slot of the compact assembly
    if (!ws->fp_dev) WS_TRY(cudaMalloc(&ws->fp_dev, 2 * sizeof(unsigned long long)));
    WS_TRY(cudaMemsetAsync(ws->fp_dev, 0, 2 * sizeof(unsigned long long), st));
    k_fingerprint<<<grid_for((long long)m + 1 + 2 * nnz, 256, 148 * 8), 256, 0, st>>>(m, nnz, ws->csr_offs, ws->csr_inds,
                                                                                      ws->csr_vals, ws->fp_dev);
    ++g_launch_count;
    WS_TRY(cudaMemcpyAsync(ws->fp, ws->fp_dev, sizeof ws->fp, cudaMemcpyDeviceToHost, st));
    WS_TRY(cudaStreamSynchronize(st));
    ws->fp_valid = true;
    ws->fp_strategy_hint = strategy_hint;
    ws->loaded = true;
    return SB200_OK;
}

int sb200_set_node_delta(sb200_ws *ws, const sb200_node_delta *delta)
{
    if (!ws) return SB200_ERR_INVALID;
    return apply_node_delta(ws, delta);
}

int sb200_solve(sb200_ws *ws, const sb200_params *params, sb200_result *result)
{
    if (!ws || !params || !result) return SB200_ERR_INVALID;
    int rc = solve_begin(ws, params, result);
    if (rc) return rc;
    int finished = 0;
    while (!finished)
    {
        if ((rc = solve_step(ws))) return rc;
        if ((rc = solve_poll(ws, &finished))) return rc;
    }
    return solve_finish(ws, result);
}

// an error in one slot must not leave the other slots of a batch marked "solve in flight"
static int abort_batch(sb200_ws **wss, int k, int rc)
{
    for (int i = 0; i < k; ++i)
        if (wss[i] && (wss[i]->active || wss[i]->window_pending))
        {
            cudaSetDevice(wss[i]->device);
            cudaStreamSynchronize(wss[i]->stream);
            wss[i]->active = false;
            wss[i]->window_pending = false;
            wss[i]->cta_launched = false;
        }
    return rc;
}

// the node deltas of a whole window: host bookkeeping per slot, then ONE staged copy and ONE kernel on the lead's stream
// (the window's launch follows on the same stream) instead of four copies and three kernels per slot
static bool deltas_can_batch(sb200_ws **wss, int k)
{
    if (k < 2) return false;
    const char *off = getenv("SB200_BATCHED_DELTAS");
    if (off && off[0] == '0') return false;
    for (int i = 0; i < k; ++i)
        if (!wss[i] || !wss[i]->loaded || wss[i]->device != wss[0]->device || wss[i]->strategy != SB200_STRATEGY_CHOLESKY)
            return false;
    return true;
}

static int apply_node_deltas_batched(sb200_ws **wss, int k, const sb200_node_delta *deltas)
{
    sb200_ws *lead = wss[0];
    sb200_ws *ws = lead;                    // for WS_TRY
    std::vector<NodeDeltaHost> hd((size_t)k);
    size_t rows = 0;
    int rc = SB200_OK, live = 0;
    for (int i = 0; i < k; ++i)
    {
        if ((rc = apply_node_delta(wss[i], &deltas[i], &hd[i])))
        {   // a rejected delta: the slots before it have changed their dimensions already, so their data still goes out
            k = i;
            break;
        }
        if (hd[i].d.k >= 0) ++live;
        if (hd[i].d.k > 0) rows += (size_t)hd[i].d.k;
    }
    const int rc_slot = rc;
    if (!live) return rc_slot;
    WS_TRY(cudaSetDevice(lead->device));
    const size_t rows_pad = (rows + 1) & ~(size_t)1;                 // the doubles behind the ids stay 8-byte aligned
    const size_t off_var = sizeof(NodeDeltaDesc) * (size_t)k, off_coef = off_var + 4 * rows_pad,
                 off_rhs = off_coef + 8 * rows_pad, bytes = off_rhs + 8 * rows_pad;
    if (bytes > lead->nd_cap)
    {
        if (lead->nd_host) cudaFreeHost(lead->nd_host);
        if (lead->nd_dev) cudaFree(lead->nd_dev);
        lead->nd_host = lead->nd_dev = nullptr;
        lead->nd_cap = 0;
        const size_t cap = bytes + bytes / 2 + 4096;
        WS_TRY(cudaMallocHost(&lead->nd_host, cap));
        WS_TRY(cudaMalloc(&lead->nd_dev, cap));
        lead->nd_cap = cap;
    }
    // (the previous window's copy out of this staging completed: every window ends with a stream synchronisation)
    NodeDeltaDesc *hdesc = reinterpret_cast<NodeDeltaDesc *>(lead->nd_host);
    int *hv = reinterpret_cast<int *>(lead->nd_host + off_var);
    double *hc = reinterpret_cast<double *>(lead->nd_host + off_coef), *hr = reinterpret_cast<double *>(lead->nd_host + off_rhs);
    size_t o = 0;
    for (int i = 0; i < k; ++i)
    {
        NodeDeltaDesc D = hd[i].d;
        if (D.k > 0)
        {
            for (int r = 0; r < D.k; ++r)
            {
                hv[o + r] = hd[i].var[r];
                hc[o + r] = hd[i].coef[r];
                hr[o + r] = hd[i].rhs[r];
            }
            D.s_var = reinterpret_cast<const int *>(lead->nd_dev + off_var) + o;
            D.s_coef = reinterpret_cast<const double *>(lead->nd_dev + off_coef) + o;
            D.s_rhs = reinterpret_cast<const double *>(lead->nd_dev + off_rhs) + o;
            o += (size_t)D.k;
        }
        hdesc[i] = D;
    }
    cudaStream_t main = lead->stream;
    WS_TRY(cudaMemcpyAsync(lead->nd_dev, lead->nd_host, bytes, cudaMemcpyHostToDevice, main));
    k_node_delta_batch<<<dim3(16, (unsigned)k), 256, 0, main>>>(reinterpret_cast<const NodeDeltaDesc *>(lead->nd_dev));
    WS_TRY(cudaGetLastError());
    ++g_launch_count;
    return rc_slot;
}

// A window of LPs in the throughput form as ONE launch: block i of k_ipm_cta solves the LP resident in wss[i].  The first
// slot's stream carries the window's deltas (or waits for the slots' own streams where they were applied one by one), the
// one LP launch and the result copies, and is the only thing the host waits on.  No limit of 128 concurrent kernels, no
// launch per LP.
static bool batch_is_one_launch(sb200_ws **wss, int k, const sb200_result *results)
{
    if (k < 2) return false;
    const char *off = getenv("SB200_WINDOW_LAUNCH");
    if (off && off[0] == '0') return false;
    for (int i = 0; i < k; ++i)
    {
        const sb200_ws *w = wss[i];
        if (!w || !w->loaded || w->device != wss[0]->device || w->solver_form != SB200_FORM_THROUGHPUT || !cta_eligible(w))
            return false;
        const sb200_result &r = results[i];
        if (r.x0_host || r.y0_host || r.s0_host) return false;
    }
    return true;
}

static int window_launch(sb200_ws **wss, int k, const sb200_params *params, sb200_result *results,
                         bool deltas_on_lead_stream)
{
    sb200_ws *lead = wss[0];
    sb200_ws *ws = lead;                    // for WS_TRY
    WS_TRY(cudaSetDevice(lead->device));
    cudaStream_t main = lead->stream;
    if (k > lead->batch_cap)
    {
        if (lead->batch_dev) cudaFree(lead->batch_dev);
        if (lead->batch_host) cudaFreeHost(lead->batch_host);
        if (lead->hbatch_dev) cudaFree(lead->hbatch_dev);
        if (lead->hbatch_host) cudaFreeHost(lead->hbatch_host);
        lead->batch_dev = nullptr; lead->batch_host = nullptr; lead->hbatch_dev = nullptr; lead->hbatch_host = nullptr;
        lead->batch_cap = 0;
        WS_TRY(cudaMalloc(&lead->batch_dev, sizeof(CtaLp) * (size_t)k));
        WS_TRY(cudaMallocHost(&lead->batch_host, sizeof(CtaLp) * (size_t)k));
        WS_TRY(cudaMalloc(&lead->hbatch_dev, sizeof(HeurArgs) * (size_t)k));
        WS_TRY(cudaMallocHost(&lead->hbatch_host, sizeof(HeurArgs) * (size_t)k));
        lead->batch_cap = k;
    }
    // parameters: one device copy (the lead's) serves every block
    DevParams &hp = lead->hparams;
    hp.eta = params->eta;
    hp.mu_tol = params->mu_tol;
    hp.min_improv_ratio = params->gap_min_improv_pct / 100.0;
    hp.max_iter = params->max_iter;
    hp.gap_enabled = (params->gap_enabled && params->gap_window > 0 && params->gap_min_improv_pct >= 0.0) ? 1 : 0;
    hp.gap_window = params->gap_window;
    hp.n_orig = lead->n_orig;
    hp.cg_max_iter = params->cg_max_iter;
    hp.cg_tol_initial = params->cg_tol_initial;
    hp.cg_tol_final = params->cg_tol_final;
    hp.cg_tol_decay = params->cg_tol_decay;
    WS_TRY(cudaMemcpyAsync(lead->dparams, &hp, sizeof hp, cudaMemcpyHostToDevice, main));
    for (int i = 0; i < k; ++i)
    {
        sb200_ws *w = wss[i];
        w->params = *params;
        w->launches_at_begin = g_launch_count;
        w->graph_kernel_launches = 0;
        fill_cta_args(w, &results[i], lead->batch_host[i]);
        lead->batch_host[i].P = lead->dparams;
        w->sc_host->done = 0;
        w->cta_launched = true;
        w->active = true;
        if (i && !deltas_on_lead_stream)
        {   // the slot's node-delta kernels (its own stream) before the window's launch
            WS_TRY(cudaEventRecord(w->ev[0], w->stream));
            WS_TRY(cudaStreamWaitEvent(main, w->ev[0], 0));
        }
    }
    WS_TRY(cudaMemcpyAsync(lead->batch_dev, lead->batch_host, sizeof(CtaLp) * (size_t)k, cudaMemcpyHostToDevice, main));
    WS_TRY(cudaEventRecord(lead->ev[1], main));
    const int rc = launch_ipm_cta(lead->batch_dev, k, main);
    if (rc) return fail(lead, rc, "sb200_solve_batch: launch of the one-block solver failed");
    WS_TRY(cudaGetLastError());
    WS_TRY(cudaEventRecord(lead->ev[2], main));
    bool copies = false;
    for (int i = 0; i < k; ++i)
    {   // results the caller wants on the host: behind the window's launch, on the same stream
        const sb200_ws *w = wss[i];
        sb200_result &r = results[i];
        if (r.x_host) WS_TRY(cudaMemcpyAsync(r.x_host, w->V.x, sizeof(double) * w->n, cudaMemcpyDeviceToHost, main));
        if (r.y_host) WS_TRY(cudaMemcpyAsync(r.y_host, w->V.y, sizeof(double) * w->m, cudaMemcpyDeviceToHost, main));
        if (r.s_host) WS_TRY(cudaMemcpyAsync(r.s_host, w->V.s, sizeof(double) * w->n, cudaMemcpyDeviceToHost, main));
        copies = copies || r.x_host || r.y_host || r.s_host;
    }
    (void)copies;
    lead->window_lps = k;
    lead->window_pending = true;
    return SB200_OK;
}

// wait for the window launched on wss[0]'s stream (and whatever was queued behind it) and collect the results
static int window_collect(sb200_ws **wss, int k, sb200_result *results)
{
    sb200_ws *lead = wss[0];
    sb200_ws *ws = lead;                    // for WS_TRY
    WS_TRY(cudaSetDevice(lead->device));
    WS_TRY(cudaStreamSynchronize(lead->stream));
    float wms = 0.f;
    cudaEventElapsedTime(&wms, lead->ev[1], lead->ev[2]);
    lead->window_ms = wms;
    lead->window_pending = false;
    for (int i = 0; i < k; ++i)
    {
        finish_from_mirror(wss[i], &results[i]);
        results[i].kernels_launched = i == 0 ? 1 : 0;        // one launch for the whole window
    }
    return SB200_OK;
}

static int solve_batch_one_launch(sb200_ws **wss, int k, const sb200_params *params, sb200_result *results,
                                  bool deltas_on_lead_stream)
{
    const int rc = window_launch(wss, k, params, results, deltas_on_lead_stream);
    return rc ? rc : window_collect(wss, k, results);
}

int sb200_solve_batch(sb200_ws **wss, int k, const sb200_node_delta *deltas, const sb200_params *params,
                      sb200_result *results)
{
    // LPs of a batch are independent: every workspace runs on its own stream and the
    // host interleaves enqueue/poll so their kernels overlap on the device.
    if (!wss || k <= 0 || !params || !results) return SB200_ERR_INVALID;
    for (int i = 0; i < k; ++i)
        if (wss[i] && wss[i]->window_pending)
            return fail(wss[i], SB200_ERR_INVALID, "sb200_solve_batch: a window begun on these workspaces has not been finished");
    std::vector<int> live(k, 0);
    int rc, remaining = 0;
    bool deltas_on_lead_stream = false;
    if (deltas)
    {
        if (deltas_can_batch(wss, k))
        {
            if ((rc = apply_node_deltas_batched(wss, k, deltas))) return abort_batch(wss, k, rc);
            deltas_on_lead_stream = true;
        }
        else
            for (int i = 0; i < k; ++i)
                if ((rc = apply_node_delta(wss[i], &deltas[i]))) return abort_batch(wss, k, rc);
    }
    for (int i = 0; i < k; ++i)
        if (deltas && deltas[i].export_xys && !results[i].xys_device) results[i].xys_device = deltas[i].export_xys;
    const bool one_launch = batch_is_one_launch(wss, k, results);
    if (deltas_on_lead_stream && !one_launch && cudaStreamSynchronize(wss[0]->stream) != cudaSuccess)
        return abort_batch(wss, k, SB200_ERR_CUDA);          // the per-slot streams below must see the applied deltas
    if (one_launch)
    {
        if ((rc = solve_batch_one_launch(wss, k, params, results, deltas_on_lead_stream))) return abort_batch(wss, k, rc);
        return SB200_OK;
    }
    for (int i = 0; i < k; ++i)
    {
        if ((rc = solve_begin(wss[i], params, &results[i]))) return abort_batch(wss, k, rc);
        live[i] = 1;
        ++remaining;
    }
    while (remaining)
    {
        for (int i = 0; i < k; ++i)
            if (live[i] && (rc = solve_step(wss[i]))) return abort_batch(wss, k, rc);
        for (int i = 0; i < k; ++i)
            if (live[i])
            {
                int fin = 0;
                if ((rc = solve_poll(wss[i], &fin))) return abort_batch(wss, k, rc);
                if (fin)
                {
                    if ((rc = solve_finish(wss[i], &results[i]))) return abort_batch(wss, k, rc);
                    live[i] = 0;
                    --remaining;
                }
            }
    }
    return SB200_OK;
}

// launch the per-node branching / incumbent kernel behind the workspace's last solve and request its 40-byte
// record (pinned); the caller waits on the stream (or an event) before reading ws->heur_out_host
static int ensure_heur_buffers(sb200_ws *ws)
{
    const int n0 = ws->n_orig;
    if (n0 > ws->heur_cap)
    {
        int rc;
        if ((rc = grow(ws, &ws->heur_list, (size_t)n0))) return rc;
        if ((rc = grow(ws, &ws->heur_sorted, (size_t)n0))) return rc;
        if ((rc = grow(ws, &ws->heur_cover, (size_t)n0))) return rc;
        if ((rc = grow(ws, &ws->heur_nif, (size_t)n0))) return rc;
        if ((rc = grow(ws, &ws->heur_score, (size_t)n0))) return rc;
        if (!ws->heur_out)
        {
            if ((rc = grow(ws, &ws->heur_out, 1))) return rc;
            WS_TRY(cudaMallocHost(&ws->heur_out_host, sizeof(sb200_heur_result)));
            WS_TRY(cudaMallocHost(&ws->heur_flag_host, sizeof(int)));
            *ws->heur_flag_host = 0;
        }
        ws->heur_cap = n0;
    }
    return SB200_OK;
}

static HeurArgs heur_args_of(sb200_ws *ws)
{
    const int n0 = ws->n_orig;
    return HeurArgs{ws->base_m, n0, ws->csr_offs, ws->csr_inds, ws->csc_colptr, ws->csc_rows, ws->c, ws->V.x,
               ws->node_k, ws->d_var, ws->d_coef, ws->heur_list, ws->heur_sorted, ws->heur_cover, ws->heur_out_host,
               ws->heur_rules, ws->heur_branch_rule, ws->heur_tol, ws->V.y, ws->b, ws->csr_vals, ws->csc_vals, ws->heur_nif, ws->heur_score, ws->heur_flag_host, ++ws->heur_seq};
}

static int enqueue_node_heuristics(sb200_ws *ws)
{
    if (!ws || !ws->loaded) return SB200_ERR_INVALID;
    if (ws->active) return fail(ws, SB200_ERR_INVALID, "sb200_node_heuristics: a solve is still in flight");
    if (!ws->csr_offs || !ws->csc_colptr || ws->n_orig <= 0)
        return fail(ws, SB200_ERR_UNSUPPORTED, "sb200_node_heuristics: the model keeps no CSR/CSC lists");
    WS_TRY(cudaSetDevice(ws->device));
    int rc0 = ensure_heur_buffers(ws);
    if (rc0) return rc0;
    const HeurArgs a = heur_args_of(ws);
    const int rc = launch_node_heuristics(a, ws->stream);
    if (rc == SB200_ERR_UNSUPPORTED)
        return fail(ws, rc, "sb200_node_heuristics: m + n_orig too large for the single-CTA kernel's shared memory");
    if (rc) return fail(ws, rc, "sb200_node_heuristics: launch configuration failed");
    WS_TRY(cudaGetLastError());      // the kernel writes its record straight into the pinned heur_out_host, then heur_flag_host
    return SB200_OK;
}

// the node rules of a whole window as ONE launch (a block per node) on wss[0]'s stream, behind whatever is queued there
static int enqueue_heuristics_window(sb200_ws **wss, int k)
{
    sb200_ws *ws = wss[0];
    int rc;
    for (int i = 0; i < k; ++i)
    {
        if ((rc = ensure_heur_buffers(wss[i]))) return rc;
        ws->hbatch_host[i] = heur_args_of(wss[i]);
        ws->hbatch_host[i].host_flag = nullptr;          // the host waits on the stream
    }
    WS_TRY(cudaMemcpyAsync(ws->hbatch_dev, ws->hbatch_host, sizeof(HeurArgs) * (size_t)k, cudaMemcpyHostToDevice, ws->stream));
    rc = launch_node_heuristics_batch(ws->hbatch_dev, k, ws->base_m, ws->n_orig, ws->stream);
    if (rc) return fail(ws, rc, "sb200_node_heuristics: window launch failed");
    WS_TRY(cudaGetLastError());
    return SB200_OK;
}
static bool heuristics_window_ok(sb200_ws **wss, int k, bool allow_active)
{
    bool one = k >= 2 && wss[0] && wss[0]->batch_cap >= k;
    for (int i = 0; one && i < k; ++i)
        one = wss[i] && wss[i]->loaded && (allow_active || !wss[i]->active) && wss[i]->device == wss[0]->device &&
              wss[i]->heur_rules == SB200_HEUR_REFERENCE && wss[i]->base_m == wss[0]->base_m &&
              wss[i]->n_orig == wss[0]->n_orig && wss[i]->csr_offs && wss[i]->csc_colptr;
    return one;
}

int sb200_node_heuristics(sb200_ws **wss, int k, sb200_heur_result *out)
{
    if (!wss || k <= 0 || !out) return SB200_ERR_INVALID;
    int rc;
    {   // a window of nodes of the same base model with the reference's rules: ONE launch, a block per node
        if (heuristics_window_ok(wss, k, false))
        {
            sb200_ws *ws = wss[0];
            WS_TRY(cudaSetDevice(ws->device));
            if ((rc = enqueue_heuristics_window(wss, k))) return rc;
            WS_TRY(cudaStreamSynchronize(ws->stream));
            for (int i = 0; i < k; ++i) out[i] = *wss[i]->heur_out_host;
            return SB200_OK;
        }
    }
    for (int i = 0; i < k; ++i)
        if ((rc = enqueue_node_heuristics(wss[i]))) return rc;
    for (int i = 0; i < k; ++i)
    {
        sb200_ws *ws = wss[i];
        WS_TRY(cudaStreamSynchronize(ws->stream));
        out[i] = *ws->heur_out_host;
    }
    return SB200_OK;
}

int sb200_solve_stream(sb200_ws **wss, int k, const sb200_params *params, sb200_next_node_fn next,
                       sb200_node_done_fn done, void *user)
{
    // Continuous batching of B&B node LPs: a slot that finishes its node takes the next one at once instead of
    // waiting for the slowest LP of a window (node LPs of one window differ by 2x in iterations).
    if (!wss || k <= 0 || !params || !next || !done) return SB200_ERR_INVALID;
    enum { IDLE = 0, SOLVING = 1, HEUR = 2 };
    std::vector<int> state(k, IDLE);
    std::vector<sb200_result> res(k);
    int rc, busy = 0;
    bool dry = false;                      // the last sweep found no node for an idle slot
    for (;;)
    {
        bool progressed = false;
        for (int i = 0; i < k; ++i)
        {
            sb200_ws *ws = wss[i];
            if (state[i] == IDLE)
            {
                if (dry && busy) continue;             // ask again only after some node has been handed back
                sb200_node_delta d{};
                if (!next(user, i, &d))
                {
                    dry = true;
                    continue;
                }
                if ((rc = apply_node_delta(ws, &d))) return abort_batch(wss, k, rc);
                res[i] = sb200_result{};
                res[i].xys_device = d.export_xys;
                if ((rc = solve_begin(ws, params, &res[i]))) return abort_batch(wss, k, rc);
                if ((rc = solve_step(ws))) return abort_batch(wss, k, rc);
                state[i] = SOLVING;
                ++busy;
                progressed = true;
            }
            else if (state[i] == SOLVING)
            {
                if (ws->cta_launched)
                {   // one-block LP: the kernel sets the pinned mirror's `done` last - a host memory read, no driver call
                    if (!*reinterpret_cast<volatile int *>(&ws->sc_host->done)) continue;
                }
                else
                {
                    const cudaError_t q = cudaEventQuery(ws->ev[3]);
                    if (q == cudaErrorNotReady) continue;
                    WS_TRY(q);
                }
                int fin = 0;
                if ((rc = solve_poll(ws, &fin))) return abort_batch(wss, k, rc);
                if (!fin)
                {
                    if ((rc = solve_step(ws))) return abort_batch(wss, k, rc);
                }
                else
                {
                    if ((rc = solve_finish(ws, &res[i]))) return abort_batch(wss, k, rc);
                    if ((rc = enqueue_node_heuristics(ws))) return abort_batch(wss, k, rc);
                    state[i] = HEUR;
                }
                progressed = true;
            }
            else
            {
                if (*reinterpret_cast<volatile int *>(ws->heur_flag_host) != ws->heur_seq) continue;    // the node kernel's flag
                done(user, i, &res[i], ws->heur_out_host);     // may add nodes to the caller's frontier
                state[i] = IDLE;
                --busy;
                dry = false;
                progressed = true;
            }
        }
        if (busy == 0 && dry) break;
        if (!progressed && busy) std::this_thread::yield();
    }
    return SB200_OK;
}

int sb200_set_concurrency_hint(sb200_ws *ws, int concurrent_lps)
{
    if (!ws || concurrent_lps < 0) return SB200_ERR_INVALID;
    int limit = 0;
    if (concurrent_lps > 1)
    {
        // CTA slots shared out: 2 per SM by default (SB200_SHARE_FACTOR overrides, for experiments)
        const char *fs = getenv("SB200_SHARE_FACTOR");
        const int factor = fs ? std::max(1, atoi(fs)) : 2;
        const int slots = factor * (ws->chol.sms > 0 ? ws->chol.sms : 148);
        const char *ms = getenv("SB200_MIN_CTAS");
        limit = std::max(ms ? std::max(1, atoi(ms)) : 4, slots / concurrent_lps);
    }
    if (limit != ws->chol.grid_limit)
    {
        WS_TRY(cudaSetDevice(ws->device));
        WS_TRY(cudaStreamSynchronize(ws->stream));
        drop_graphs(ws);                       // the launch geometry is part of the captured iteration
        ws->chol.grid_limit = limit;
    }
    // many LPs in flight: one CTA per LP (sb200_cta.cu) unless SB200_CTA_SOLVER=0 asks for the shared multi-kernel form
    const char *cs = getenv("SB200_CTA_SOLVER");
    ws->solver_form = (concurrent_lps > 1 && !(cs && atoi(cs) == 0)) ? SB200_FORM_THROUGHPUT : SB200_FORM_LATENCY;
    return SB200_OK;
}

int sb200_window_begin(sb200_ws **wss, int k, const sb200_node_delta *deltas, const sb200_params *params,
                       sb200_result *results, int with_node_rules)
{
    if (!wss || k <= 0 || !params || !results) return SB200_ERR_INVALID;
    for (int i = 0; i < k; ++i)
        if (!wss[i] || wss[i]->active || wss[i]->window_pending) return SB200_ERR_INVALID;
    int rc;
    // only what the one-launch window can run is accepted here; nothing has changed when it is not
    {
        const char *off = getenv("SB200_WINDOW_LAUNCH");
        if ((off && off[0] == '0') || k < 2 || (deltas && !deltas_can_batch(wss, k))) return SB200_ERR_UNSUPPORTED;
        for (int i = 0; i < k; ++i)
            if (!wss[i]->loaded || wss[i]->device != wss[0]->device || wss[i]->solver_form != SB200_FORM_THROUGHPUT ||
                !cta_eligible(wss[i]) || results[i].x0_host || results[i].y0_host || results[i].s0_host)
                return SB200_ERR_UNSUPPORTED;
    }
    if (deltas && (rc = apply_node_deltas_batched(wss, k, deltas))) return abort_batch(wss, k, rc);
    for (int i = 0; i < k; ++i)
        if (deltas && deltas[i].export_xys && !results[i].xys_device) results[i].xys_device = deltas[i].export_xys;
    if (!batch_is_one_launch(wss, k, results))
    {   // (a node deeper than the one-block solver takes: the deltas are applied, the caller solves with sb200_solve_batch
        // and deltas = NULL)
        cudaStreamSynchronize(wss[0]->stream);
        return SB200_ERR_UNSUPPORTED;
    }
    if ((rc = window_launch(wss, k, params, results, deltas != nullptr))) return abort_batch(wss, k, rc);
    if (with_node_rules)
    {
        if (!heuristics_window_ok(wss, k, true))
        {
            window_collect(wss, k, results);
            return fail(wss[0], SB200_ERR_INVALID, "sb200_window_begin: the node rules need sb200_set_heuristic_rules(REFERENCE) on every slot");
        }
        if ((rc = enqueue_heuristics_window(wss, k)))
        {
            window_collect(wss, k, results);
            return rc;
        }
    }
    return SB200_OK;
}

int sb200_window_finish(sb200_ws **wss, int k, sb200_result *results, sb200_heur_result *rules_out)
{
    if (!wss || k <= 0 || !results || !wss[0] || !wss[0]->window_pending || wss[0]->window_lps != k) return SB200_ERR_INVALID;
    const int rc = window_collect(wss, k, results);
    if (rc) return abort_batch(wss, k, rc);
    if (rules_out)
        for (int i = 0; i < k; ++i) rules_out[i] = *wss[i]->heur_out_host;
    return SB200_OK;
}

int sb200_prepare_nodes(sb200_ws *ws, int max_extra_rows)
{
    if (!ws || !ws->loaded || max_extra_rows < 0) return SB200_ERR_INVALID;
    if (ws->strategy != SB200_STRATEGY_CHOLESKY)
        return fail(ws, SB200_ERR_UNSUPPORTED, "sb200_prepare_nodes: only the sparse-assembly + Cholesky strategy folds node rows");
    if (ws->node_k != 0) return fail(ws, SB200_ERR_INVALID, "sb200_prepare_nodes: call it on the base model");
    WS_TRY(cudaSetDevice(ws->device));
    int rc;
    if ((rc = ensure_node_buffers(ws, max_extra_rows > 0 ? max_extra_rows : 1, true))) return rc;
    if ((rc = ensure_heur_buffers(ws))) return rc;
    WS_TRY(cudaStreamSynchronize(ws->stream));
    return SB200_OK;
}

int sb200_last_window(sb200_ws *ws, double *ms, int *lps)
{
    if (!ws) return SB200_ERR_INVALID;
    if (ms) *ms = ws->window_ms;
    if (lps) *lps = ws->window_lps;
    return SB200_OK;
}

int sb200_set_solver_form(sb200_ws *ws, int form)
{
    if (!ws || (form != SB200_FORM_LATENCY && form != SB200_FORM_THROUGHPUT)) return SB200_ERR_INVALID;
    ws->solver_form = form;
    return SB200_OK;
}

int sb200_get_cover(sb200_ws *ws, unsigned char *x_host)
{
    if (!ws || !ws->loaded || !x_host || !ws->heur_cover) return SB200_ERR_INVALID;
    WS_TRY(cudaSetDevice(ws->device));
    WS_TRY(cudaMemcpyAsync(x_host, ws->heur_cover, (size_t)ws->n_orig, cudaMemcpyDeviceToHost, ws->stream));
    WS_TRY(cudaStreamSynchronize(ws->stream));
    return SB200_OK;
}

int sb200_get_rounded(sb200_ws *ws, unsigned char *x_host)
{
    if (!ws || !ws->loaded || !x_host || !ws->heur_nif) return SB200_ERR_INVALID;
    WS_TRY(cudaSetDevice(ws->device));
    WS_TRY(cudaMemcpyAsync(x_host, ws->heur_nif, (size_t)ws->n_orig, cudaMemcpyDeviceToHost, ws->stream));
    WS_TRY(cudaStreamSynchronize(ws->stream));
    return SB200_OK;
}

int sb200_set_heuristic_rules(sb200_ws *ws, int rules, int branch_rule, double integrality_tol)
{
    if (!ws || (rules != SB200_HEUR_PLAIN && rules != SB200_HEUR_REFERENCE) ||
        (branch_rule != SB200_BRANCH_MOST_FRACTIONAL && branch_rule != SB200_BRANCH_HIGHEST_COST_FRACTIONAL) ||
        !(integrality_tol >= 0.0))
        return SB200_ERR_INVALID;
    ws->heur_rules = rules;
    ws->heur_branch_rule = branch_rule;
    ws->heur_tol = integrality_tol;
    return SB200_OK;
}

int sb200_get_trace(sb200_ws *ws, double *out, int max_rows)
{
    if (!ws || !out) return 0;
    const int rows = std::min(max_rows, ws->trace_rows);
    if (ws->trace_stale)
    {   // one-block solves leave the trace on the device until somebody asks for it
        ws->trace_host.resize((size_t)ws->trace_rows * SB200_TRACE_COLS);
        if (ws->trace_rows)
        {
            cudaSetDevice(ws->device);
            cudaStreamSynchronize(ws->stream);
            cudaMemcpy(ws->trace_host.data(), ws->V.trace, sizeof(double) * ws->trace_host.size(), cudaMemcpyDeviceToHost);
        }
        ws->trace_stale = false;
    }
    if (rows > 0) memcpy(out, ws->trace_host.data(), sizeof(double) * (size_t)rows * SB200_TRACE_COLS);
    return rows;
}

int sb200_get_device_iterates(sb200_ws *ws, void **x, void **y, void **s)
{
    if (!ws || !ws->loaded) return SB200_ERR_INVALID;
    if (x) *x = ws->V.x;
    if (y) *y = ws->V.y;
    if (s) *s = ws->V.s;
    return SB200_OK;
}

int sb200_get_iterates(sb200_ws *ws, double *x_host, double *y_host, double *s_host)
{
    if (!ws || !ws->loaded) return SB200_ERR_INVALID;
    WS_TRY(cudaSetDevice(ws->device));
    if (x_host) WS_TRY(cudaMemcpyAsync(x_host, ws->V.x, sizeof(double) * ws->n, cudaMemcpyDeviceToHost, ws->stream));
    if (y_host) WS_TRY(cudaMemcpyAsync(y_host, ws->V.y, sizeof(double) * ws->m, cudaMemcpyDeviceToHost, ws->stream));
    if (s_host) WS_TRY(cudaMemcpyAsync(s_host, ws->V.s, sizeof(double) * ws->n, cudaMemcpyDeviceToHost, ws->stream));
    WS_TRY(cudaStreamSynchronize(ws->stream));
    return SB200_OK;
}

int sb200_model_info(sb200_ws *ws, long long *info, int n_info)
{
    if (!ws || !ws->loaded || !info) return SB200_ERR_INVALID;
    const long long vals[] = {ws->m, ws->n, ws->n_orig, ws->nnz, ws->mpad, ws->strategy, ws->pat.n_pairs,
                              ws->pat.n_terms, ws->pat.term_w ? 1 : 0, ws->csc_lanes,
                              ws->iter_graph_kernels, ws->chol.max_coop_grid, ws->blk_rows.ptr ? 1 : 0,
                              ws->blk_rows.nblk, ws->blk_cols.nblk, (long long)ws->blk_rows.n_chunks,
                              (long long)ws->blk_cols.n_chunks, (long long)ws->pat.n_chunks};
    const int k = (int)(sizeof vals / sizeof vals[0]);
    for (int i = 0; i < n_info && i < k; ++i) info[i] = vals[i];
    return k;
}

int sb200_time_phase(sb200_ws *ws, int phase, int reps, double *ms_out)
{
    if (!ws || !ws->loaded || !ms_out || reps <= 0) return SB200_ERR_INVALID;
    if (phase >= 100 && phase < 100 + 2 * SB200_TRACE_COLS)
    {   // 108, 109, 110: the factorisation split into accumulation steps, diagonal tiles, chunk epilogues   // phases of the LAST one-CTA solve on this workspace (sb200_cta.cu leaves them in the last trace row, ns):
        // 100 assembly, 101 factorisation, 102 solves, 103 A v, 104 A' v + epilogues, 105 vector steps,
        // 106 starting point, 107 whole LP
        double ns = 0.0;
        WS_TRY(cudaSetDevice(ws->device));
        WS_TRY(cudaStreamSynchronize(ws->stream));
        WS_TRY(cudaMemcpy(&ns, ws->V.trace + (size_t)(SB200_TRACE_ROWS - 1) * SB200_TRACE_COLS +
                                   (phase < 108 ? phase - 100 : phase - 108 - SB200_TRACE_COLS), sizeof(double),
                          cudaMemcpyDeviceToHost));
        *ms_out = ns * 1e-6;
        return SB200_OK;
    }
    if (ws->strategy == SB200_STRATEGY_PCG && (phase == 0 || phase == 1 || phase == 2))
        return fail(ws, SB200_ERR_UNSUPPORTED, "sb200_time_phase: direct-path phase on a PCG model");
    WS_TRY(cudaSetDevice(ws->device));
    cudaStream_t st = ws->stream;
    const IpmVecs &V = ws->V;
    cudaEvent_t e0, e1;
    WS_TRY(cudaEventCreate(&e0));
    WS_TRY(cudaEventCreate(&e1));
    // a benign state: d = 1, done = 0
    launch_reset_scalars(ws->sc, st);
    if (phase >= 6 && phase <= 8 && ws->strategy == SB200_STRATEGY_PCG)
    {   // CG state for rhs = b, D = I, tolerance never met
        PcgVecs C{ws->m, V.rhs, ws->cg_diag, ws->cg_x, ws->cg_r, ws->cg_z, ws->cg_p, ws->cg_Ap, ws->cg_q, ws->ones_n,
                  V.partial};
        WS_TRY(cudaMemsetAsync(V.rhs, 0, sizeof(double) * ws->mpad, st));
        WS_TRY(cudaMemcpyAsync(V.rhs, ws->b, sizeof(double) * ws->m, cudaMemcpyDeviceToDevice, st));
        launch_jacobi_diag(csr_of(ws), ws->ones_n, ws->cg_diag, st);
        launch_cg_init(C, ws->sc, ws->dparams, 1e-300, 0, st);
    }
    double total = 0.0;
    for (int r = -1; r < reps; ++r)
    {
        if (phase == 1 || phase == 2)
        {   // the factorisation is in place: re-assemble (untimed) before every timed potrf
            if (ws->strategy == SB200_STRATEGY_SYRK)
            {
                launch_syrk_dmma(ws->m, ws->n, ws->denseA, ws->kpad, ws->ones_n, ws->M, ws->mpad, st);
                launch_pad_identity(ws->m, ws->M, ws->mpad, st);
            }
            else
                launch_assemble_normal(ws->pat, ws->ones_n, ws->M, ws->mpad, st);
            if (phase == 2) launch_potrf(ws->chol, ws->m, ws->M, ws->mpad, &ws->sc->chol_info, st);
        }
        WS_TRY(cudaEventRecord(e0, st));
        switch (phase)
        {
        case 0:
            if (ws->strategy == SB200_STRATEGY_SYRK)
                launch_syrk_dmma(ws->m, ws->n, ws->denseA, ws->kpad, ws->ones_n, ws->M, ws->mpad, st);
            else
                launch_assemble_normal(ws->pat, ws->ones_n, ws->M, ws->mpad, st);
            break;
        case 1: launch_potrf(ws->chol, ws->m, ws->M, ws->mpad, &ws->sc->chol_info, st); break;
        case 2: launch_potrs(ws->chol, ws->m, ws->M, ws->mpad, V.rhs, st); break;
        case 3: launch_spmv_csr(csr_of(ws), V.t, V.resB, V.rhs, 1.0, 1.0, st); break;
        case 4: launch_spmv_csc(csc_of(ws), CSC_RECOVER, V.y, nullptr, nullptr, 0, 0, &V, st); break;
        case 5: launch_affine_corrector(V, st); break;
        case 6:
        case 7:
        case 8:
        {   // PCG strategy: one CG iteration / its A'p product / its A q product on a never-converging solve
            if (ws->strategy != SB200_STRATEGY_PCG)
                return fail(ws, SB200_ERR_UNSUPPORTED, "sb200_time_phase: CG phase on a direct-strategy model");
            PcgVecs C{ws->m, V.rhs, ws->cg_diag, ws->cg_x, ws->cg_r, ws->cg_z, ws->cg_p, ws->cg_Ap, ws->cg_q,
                      ws->ones_n, V.partial};
            if (phase == 6)
                launch_cg_iteration(csr_of(ws), csc_of(ws), C, V, ws->dparams, 1 << 30, st);
            else if (phase == 7)
                launch_spmv_csc_cg(csc_of(ws), C.p, C.q, C.dscale, ws->sc, st);
            else if (csr_of(ws).blk)
                launch_blk_cg_matvec(*csr_of(ws).blk, C, ws->sc, st);
            else
                launch_spmv_csr(csr_of(ws), C.q, nullptr, C.Ap, 1.0, 0.0, st);
            break;
        }
        default: return fail(ws, SB200_ERR_INVALID, "sb200_time_phase: unknown phase");
        }
        WS_TRY(cudaEventRecord(e1, st));
        WS_TRY(cudaEventSynchronize(e1));
        float ms = 0.f;
        WS_TRY(cudaEventElapsedTime(&ms, e0, e1));
        if (r >= 0) total += ms;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    *ms_out = total / reps;
    return SB200_OK;
}

int sb200_ws_spmv(sb200_ws *ws, int transpose, const double *d_x, double *d_y)
{
    if (!ws || !ws->loaded || !d_x || !d_y) return SB200_ERR_INVALID;
    WS_TRY(cudaSetDevice(ws->device));
    if (transpose)
        launch_spmv_csc(csc_of(ws), CSC_PLAIN, d_x, d_y, d_y, 1.0, 0.0, nullptr, ws->stream);
    else
        launch_spmv_csr(csr_of(ws), d_x, d_y, d_y, 1.0, 0.0, ws->stream);
    WS_TRY(cudaGetLastError());
    WS_TRY(cudaStreamSynchronize(ws->stream));
    return SB200_OK;
}

// ---- L0 kernels ------------------------------------------------------------------------------------
static int l0_done(cudaStream_t st, bool sync)
{
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess && sync) e = cudaStreamSynchronize(st);
    return e == cudaSuccess ? SB200_OK : SB200_ERR_CUDA;
}

int sb200_k_elem_min_mult(const double *d_x, const double *d_s, double *d_out, int n, void *stream)
{
    if (n <= 0) return SB200_OK;
    launch_elem_min_mult(d_x, d_s, d_out, n, (cudaStream_t)stream);
    return l0_done((cudaStream_t)stream, false);
}

int sb200_k_corrector_rhs(const double *d_dx, const double *d_ds, double sigma, double mu, double *d_out,
                          int n, void *stream)
{
    if (n <= 0) return SB200_OK;
    launch_corrector_rhs(d_dx, d_ds, sigma, mu, d_out, n, (cudaStream_t)stream);
    return l0_done((cudaStream_t)stream, false);
}

int sb200_k_alpha_max(const double *d_x, const double *d_dx, const double *d_s, const double *d_ds, int n,
                      double *d_result, double *h_result, void *stream)
{
    cudaStream_t st = (cudaStream_t)stream;
    unsigned long long *ord = nullptr;
    if (cudaMallocAsync(&ord, 16, st) != cudaSuccess) return SB200_ERR_NOMEM;
    launch_alpha_max(d_x, d_dx, d_s, d_ds, n > 0 ? n : 0, ord, d_result, st);
    cudaFreeAsync(ord, st);
    if (h_result)
    {
        if (cudaMemcpyAsync(h_result, d_result, 16, cudaMemcpyDeviceToHost, st) != cudaSuccess)
            return SB200_ERR_CUDA;
        return l0_done(st, true);
    }
    return l0_done(st, false);
}

int sb200_k_spmv_csr(int m, const int *d_offs, const int *d_inds, const double *d_vals, const double *d_x,
                     double *d_y, double alpha, double beta, void *stream)
{
    if (m <= 0) return SB200_OK;
    launch_spmv_csr(CsrView{m, d_offs, d_inds, d_vals}, d_x, d_y, d_y, alpha, beta, (cudaStream_t)stream);
    return l0_done((cudaStream_t)stream, false);
}

int sb200_k_spmv_csc(int n, const int *d_colptr, const int *d_rows, const double *d_vals, const double *d_x,
                     double *d_y, double alpha, double beta, void *stream)
{
    if (n <= 0) return SB200_OK;
    launch_spmv_csc(CscView{n, d_colptr, d_rows, d_vals, 8}, CSC_PLAIN, d_x, d_y, d_y, alpha, beta, nullptr,
                    (cudaStream_t)stream);
    return l0_done((cudaStream_t)stream, false);
}

int sb200_k_jacobi_diag(int m, const int *d_offs, const int *d_inds, const double *d_vals, const double *d_d,
                        double *d_diag, void *stream)
{
    if (m <= 0) return SB200_OK;
    launch_jacobi_diag(CsrView{m, d_offs, d_inds, d_vals}, d_d, d_diag, (cudaStream_t)stream);
    return l0_done((cudaStream_t)stream, false);
}

static CholWork g_l0_chol;     // scratch of the stand-alone potrf/potrs entry points
static ErrorSink g_l0_err;

int sb200_k_potrf(int n, double *d_a, int ld, int *d_info, void *stream)
{
    if (n <= 0 || ld % SB200_TILE != 0 || ld < n) return SB200_ERR_INVALID;
    cudaStream_t st = (cudaStream_t)stream;
    int rc = chol_work_ensure(g_l0_err, g_l0_chol, ld);
    if (rc) return rc;
    if (cudaMemsetAsync(d_info, 0, sizeof(int), st) != cudaSuccess) return SB200_ERR_CUDA;
    launch_potrf(g_l0_chol, n, d_a, ld, d_info, st);
    return l0_done(st, false);
}

int sb200_k_potrs(int n, const double *d_l, int ld, double *d_b, void *stream)
{
    if (n <= 0 || ld % SB200_TILE != 0 || ld < n) return SB200_ERR_INVALID;
    cudaStream_t st = (cudaStream_t)stream;
    int rc = chol_work_ensure(g_l0_err, g_l0_chol, ld);
    if (rc) return rc;
    launch_potrs(g_l0_chol, n, d_l, ld, d_b, st);
    return l0_done(st, false);
}

int sb200_k_syrk(int m, int k, const double *d_a, int lda, const double *d_d, double *d_c, int ld, void *stream)
{
    if (m <= 0 || k <= 0 || ld % SB200_TILE != 0 || lda % 32 != 0 || lda < k) return SB200_ERR_INVALID;
    launch_syrk_dmma(m, k, d_a, lda, d_d, d_c, ld, (cudaStream_t)stream);
    return l0_done((cudaStream_t)stream, false);
}

int sb200_assemble_normal(sb200_ws *ws, const double *d_d, double *d_m, int ld)
{
    if (!ws || !ws->loaded) return SB200_ERR_INVALID;
    if (ws->strategy == SB200_STRATEGY_SYRK)
    {
        if (ld != ws->mpad) return fail(ws, SB200_ERR_INVALID, "sb200_assemble_normal: ld must equal mpad");
        launch_syrk_dmma(ws->m, ws->n, ws->denseA, ws->kpad, d_d, d_m, ld, ws->stream);
    }
    else if (ws->pat.pair_ptr)
    {   // the compact term lists index one slot past the end (d[n] = 0): go through a workspace copy
        WS_TRY(cudaMemcpyAsync(ws->cg_q, d_d, sizeof(double) * ws->n, cudaMemcpyDeviceToDevice, ws->stream));
        launch_assemble_normal(ws->pat, ws->cg_q, d_m, ld, ws->stream);
    }
    else
        return fail(ws, SB200_ERR_UNSUPPORTED, "sb200_assemble_normal: model was loaded for the PCG strategy");
    return l0_done(ws->stream, true);
}

} // extern "C"
