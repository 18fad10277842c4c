// sb200_heur.cu - the per-node combinatorial step of the batched B&B, on the device.
//
// After a node's LP relaxation the reference's driver runs, on the host and one node at a time,
//   * the branching-variable rule MostFractionalSelector
//     (/root/reference/src/sypha_solver_heuristics.cpp:10-30), and
//   * its integer heuristics NearestIntegerFixingHeuristic (:53-110) and the cover repair of
//     DualGuidedCoverRepairHeuristic (:112-292)
// on a host copy of the LP point (/root/reference/src/sypha_solver_bnb_driver.cpp:861-1005).  With the LP
// itself at a few milliseconds, that host step (about 2 ms per node in NumPy) bounds nodes/s once 16+ node
// LPs run concurrently.  Here the plain versions of those rules run as ONE single-CTA
// kernel per node on the node's own stream, right behind the LP, reading the LP point where the solver left
// it; K nodes of a window overlap on the device and 48 bytes per node return to the host.
//
// What the kernel computes (identical, tie-breaks included, to sypha_b200/bnb.py::CoverHeuristic, which the
// tests keep as the checker):
//   1. x_j = [x_lp_j >= 0.5] for the original columns, columns fixed to 0 by the node's decisions excluded;
//   2. greedy repair: while a row is uncovered take the usable column with the smallest cost per newly
//      covered row (first index on ties); gains are kept up to date incrementally;
//   3. redundant columns are dropped, dearest first (ties: lowest index first);
//   4. branching variable = most fractional original column (first index on ties), |x - rint(x)|;
//   5. c . rint(x), the incumbent offer when the LP point is integral.
// All integer work is exact; the two objectives are sums of products cost x {0,1} in a fixed tree order.
#include "sb200_kernels.cuh"
#include "sb200_heur.cuh"

namespace sb200 {

namespace {

constexpr int HT = 1024;          // threads of the single CTA
constexpr int HW = HT / 32;

struct MinKey
{
    double v;
    int j;
};

__device__ __forceinline__ bool key_less(double va, int ja, double vb, int jb)
{
    return va < vb || (va == vb && ja < jb);
}

// block-wide lexicographic minimum of (v, j); result valid in every thread
__device__ MinKey block_min(double v, int j, MinKey *sred)
{
    const unsigned full = 0xffffffffu;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
    {
        const double ov = __shfl_xor_sync(full, v, o);
        const int oj = __shfl_xor_sync(full, j, o);
        if (key_less(ov, oj, v, j)) { v = ov; j = oj; }
    }
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    __syncthreads();                       // sred may still be read from the previous call
    if (l == 0) { sred[w].v = v; sred[w].j = j; }
    __syncthreads();
    v = sred[l].v;                         // HW == 32: one entry per lane
    j = sred[l].j;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
    {
        const double ov = __shfl_xor_sync(full, v, o);
        const int oj = __shfl_xor_sync(full, j, o);
        if (key_less(ov, oj, v, j)) { v = ov; j = oj; }
    }
    return MinKey{v, j};
}

__device__ double block_sum(double v, double *sred)
{
    const unsigned full = 0xffffffffu;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(full, v, o);
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    __syncthreads();
    if (l == 0) sred[w] = v;
    __syncthreads();
    v = sred[l];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(full, v, o);
    return v;
}

// state bits of a column
constexpr unsigned char ST_X = 1, ST_BANNED = 2;

__global__ void __launch_bounds__(HT, 1) k_node_heuristics(HeurArgs a)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    int *cover = reinterpret_cast<int *>(smem_raw);                 // [m0]
    int *gain = cover + a.m0;                                       // [n0]
    unsigned char *state = reinterpret_cast<unsigned char *>(gain + a.n0);   // [n0]
    __shared__ MinKey sred[HW];
    __shared__ double sredd[HW];
    __shared__ int s_unc, s_chosen, s_steps;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int m0 = a.m0, n0 = a.n0;

    // ---- 4./5. branching variable and the rounded objective (on the LP point itself) -------------------
    {
        double best = -1.0;
        int bj = n0;
        double racc = 0.0;
        for (int j = tid; j < n0; j += HT)
        {
            const double x = a.x_lp[j];
            const double r = rint(x);                               // half to even, as numpy.round
            const double f = fabs(x - r);
            if (f > best) { best = f; bj = j; }                     // j ascends: first index kept on ties
            racc += a.c[j] * r;
        }
        const MinKey k = block_min(-best, bj, sred);
        const double rsum = block_sum(racc, sredd);
        if (tid == 0)
        {
            a.out->branch_var = k.j < n0 ? k.j : -1;
            a.out->branch_frac = -k.v;
            a.out->rounded_obj = rsum;
        }
    }

    // ---- 1. rounding ---------------------------------------------------------------------------------
    for (int j = tid; j < n0; j += HT) state[j] = a.x_lp[j] >= 0.5 ? ST_X : 0;
    for (int i = tid; i < m0; i += HT) cover[i] = 0;
    if (tid == 0) { s_unc = 0; s_chosen = 0; s_steps = 0; }
    __syncthreads();
    for (int r = tid; r < a.k; r += HT)
        if (a.coef[r] < 0.0)                                        // decision "var = 0" (bnb.cpp:453-468)
        {
            const int v = a.var[r];
            if (v >= 0 && v < n0) state[v] = ST_BANNED;
        }
    __syncthreads();
    // cover_i = number of chosen columns in row i.  Through the CSC lists of the CHOSEN columns (a few dozen), not
    // through all rows of the CSR copy: one CTA reading the whole matrix twice (cover counts by rows, gains by
    // columns) was most of this kernel's 350 us.
    for (int j = tid; j < n0; j += HT)
        if (state[j] & ST_X) a.list[atomicAdd(&s_chosen, 1)] = j;
    __syncthreads();
    {
        const int nsel = s_chosen;
        for (int q = warp; q < nsel; q += HW)
        {
            const int j = a.list[q];
            for (int p = a.col_ptr[j] + lane; p < a.col_ptr[j + 1]; p += 32)
            {
                const int i = a.col_rows[p];
                if (i < m0) atomicAdd(&cover[i], 1);
            }
        }
    }
    __syncthreads();
    for (int i = tid; i < m0; i += HT)
        if (cover[i] == 0) atomicAdd(&s_unc, 1);
    if (tid == 0) s_chosen = 0;                 // reused by the redundancy pass
    __syncthreads();

    // ---- 2. greedy repair ------------------------------------------------------------------------------
    int feasible = 1;
    if (s_unc > 0)
    {
        // gain_j = uncovered rows column j would cover: every uncovered row (few) adds one to the columns it holds
        for (int j = tid; j < n0; j += HT) gain[j] = 0;
        __syncthreads();
        for (int i = warp; i < m0; i += HW)
            if (cover[i] == 0)
                for (int p = a.row_ptr[i] + lane; p < a.row_ptr[i + 1]; p += 32)
                {
                    const int j = a.row_cols[p];
                    if (j < n0) atomicAdd(&gain[j], 1);
                }
        __syncthreads();
        while (true)
        {
            if (s_unc == 0) break;                                  // uniform: read after a barrier
            double best = DBL_MAX;
            int bj = n0;
            for (int j = tid; j < n0; j += HT)
                if (state[j] == 0 && gain[j] > 0)
                {
                    const double sc = a.c[j] / (double)gain[j];
                    if (sc < best) { best = sc; bj = j; }
                }
            const MinKey k = block_min(best, bj, sred);
            if (k.j >= n0) { feasible = 0; break; }                 // no usable column covers anything new
            const int jc = k.j;
            if (tid == 0) { state[jc] = ST_X; ++s_steps; }
            // rows of the chosen column: warp per row; a newly covered row lowers the gain of its columns
            const int c0 = a.col_ptr[jc], c1 = a.col_ptr[jc + 1];
            for (int q = c0 + warp; q < c1; q += HW)
            {
                const int i = a.col_rows[q];
                if (i >= m0) continue;
                int old = 0;
                if (lane == 0) { old = cover[i]; cover[i] = old + 1; }   // a row appears once per column
                old = __shfl_sync(0xffffffffu, old, 0);
                if (old == 0)
                {
                    if (lane == 0) atomicSub(&s_unc, 1);
                    for (int p = a.row_ptr[i] + lane; p < a.row_ptr[i + 1]; p += 32)
                    {
                        const int j = a.row_cols[p];
                        if (j < n0) atomicSub(&gain[j], 1);
                    }
                }
            }
            __syncthreads();
        }
    }
    __syncthreads();

    // ---- 3. drop redundant columns, dearest first ------------------------------------------------------
    if (feasible)
    {
        for (int j = tid; j < n0; j += HT)
            if (state[j] & ST_X) a.list[atomicAdd(&s_chosen, 1)] = j;
        __syncthreads();
        const int nc = s_chosen;
        for (int p = tid; p < nc; p += HT)
        {   // rank by counting under the order (cost descending, index ascending)
            const int jp = a.list[p];
            const double cp = a.c[jp];
            int rank = 0;
            for (int q = 0; q < nc; ++q)
            {
                const int jq = a.list[q];
                const double cq = a.c[jq];
                rank += (cq > cp || (cq == cp && jq < jp)) ? 1 : 0;
            }
            a.sorted[rank] = jp;
        }
        __syncthreads();
        if (warp == 0)
        {
            for (int t = 0; t < nc; ++t)
            {
                const int j = a.sorted[t];
                const int c0 = a.col_ptr[j], c1 = a.col_ptr[j + 1];
                bool red = true;
                for (int q = c0 + lane; q < c1; q += 32)
                {
                    const int i = a.col_rows[q];
                    if (i < m0 && cover[i] < 2) red = false;
                }
                if (__all_sync(0xffffffffu, red))
                {
                    for (int q = c0 + lane; q < c1; q += 32)
                    {
                        const int i = a.col_rows[q];
                        if (i < m0) cover[i] -= 1;
                    }
                    if (lane == 0) state[j] = 0;
                }
                __syncwarp();
            }
        }
        __syncthreads();
    }

    // ---- objective and the cover itself ------------------------------------------------------------------
    double acc = 0.0;
    int cnt = 0;
    for (int j = tid; j < n0; j += HT)
    {
        const unsigned char sel = (state[j] & ST_X) ? 1 : 0;
        a.cover_x[j] = sel;
        if (sel) { acc += a.c[j]; ++cnt; }
    }
    const double obj = block_sum(acc, sredd);
    const double nsel = block_sum((double)cnt, sredd);
    if (tid == 0)
    {
        a.out->feasible = feasible;
        a.out->cover_obj = feasible ? obj : DBL_MAX;
        a.out->n_chosen = (int)nsel;
        a.out->repair_steps = s_steps;
    }
}

} // namespace

size_t heur_smem_bytes(int m0, int n0) { return sizeof(int) * ((size_t)m0 + (size_t)n0) + (size_t)n0 + 16; }

int launch_node_heuristics(const HeurArgs &a, cudaStream_t st)
{
    const size_t smem = heur_smem_bytes(a.m0, a.n0);
    if (smem > 200 * 1024) return SB200_ERR_UNSUPPORTED;
    if (smem > 48 * 1024 &&      // per device, so not cached
        cudaFuncSetAttribute(k_node_heuristics, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024) != cudaSuccess)
        return SB200_ERR_CUDA;
    k_node_heuristics<<<1, HT, smem, st>>>(a);
    ++g_launch_count;
    return SB200_OK;
}

} // namespace sb200
